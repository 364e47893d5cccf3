"""CPU checkers for the CUDA path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package; the product (cuda-recommender_b200/) never does.

  oracle.port   ctypes binding of oracle/libmforacle.so  (restatement, mf_oracle.c)
  oracle.ref    ctypes binding of oracle/_ref/libmfref.so (unmodified reference CPU
                sources compiled by oracle/Makefile; may be absent)
"""
from . import port  # noqa: F401
from . import ref  # noqa: F401
