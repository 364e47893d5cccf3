// extras.h — command line, post-run checks (reference: /root/reference/src/extras.{h,cpp}).
#ifndef B200_EXTRAS_H
#define B200_EXTRAS_H

#include "pmf.h"
#include "tools.h"

void exit_with_help();
parameter parse_command_line(int argc, char** argv);
void calculate_rmse_directly(MatData& W, MatData& H, TestData& T, int rank, bool ifALS);
void golden_compare(const MatData& W, const MatData& W_ref, unsigned k, unsigned m);

#endif  // B200_EXTRAS_H
