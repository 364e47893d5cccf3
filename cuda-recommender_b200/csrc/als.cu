// als.cu — ALS half-step: per-segment Gram assembly + Cholesky solve (sm_100a).
//
// Reference: ALS_OMP row/column bodies (src/ALS.cpp:98-158, :161-219) with Mt_byM_multiply (:66-79) and
// the explicit Cholesky inverse (:6-64); GPU kernels being replaced: updateW_overH_kernel /
// updateH_overW_kernel (cuda_src/ALS_CUDA.cu:81-181), which give every row to ONE thread and
// malloc() k*k floats per thread on the device heap.
//
// Here one CTA owns a segment (a user row or an item column) at a time — or one part of a long segment, see AlsItem —
// and takes its work from a list sorted longest-first through an atomic queue that runs one item ahead.  Everything
// about a segment is one register-tiled computation on
// the augmented matrix  M = [Y_O | r]^T [Y_O | r]  (+ lambda on the first k diagonal entries, lambda NOT scaled by
// |O|, src/ALS.cpp:120-122):
//   * the segment's factor rows Y[idx] are staged through shared memory in batches of 32 rows with asynchronous
//     copies (cp.async, three buffers, one barrier per batch, one row per lane), the rating r is staged as one more
//     column behind the k factor columns;
//   * M is accumulated in registers as TS x TS tiles of the lower triangle (TS = 8, or 4 for small k), one tile per
//     thread, the rows of a batch dealt to `ks` thread groups (split-K) whose partial tiles are added in a fixed
//     order through shared memory: the Gram matrix A = M[0:k,0:k] and the right-hand side b = M[k,0:k] come out of
//     the same loop;
//   * the tiles never leave the registers for the factorisation: blocked right-looking Cholesky (diagonal tile
//     factored by its owner, panel tiles solved against it, trailing tiles updated by the same rank-1 code as the
//     Gram loop).  Row k of the factor of M is y = L^-1 b, so the forward substitution costs nothing;
//   * L^T x = y is solved block by block by one warp — no explicit inverse;
//   * an empty segment writes zeros (src/ALS.cpp:151-157).
// Shared-memory layout of a staged row / of a column of L: 4-float chunks; with TS = 8 the two chunks of tile t
// sit at chunk positions t and nb + t, so that consecutive tiles read consecutive 16-byte words (no bank conflicts).
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "session.cuh"

namespace mf {
namespace {

constexpr int kBatch = 32;  // factor rows staged per step
constexpr int kStages = 3;  // staging buffers
// register caps of the kernel classes (A/B knobs: scripts/build_variant.sh)
#ifndef MF_ALS_REG4
#define MF_ALS_REG4 64
#endif
#ifndef MF_ALS_REG8
#define MF_ALS_REG8 168
#endif
#ifndef MF_ALS_UNROLL8
#define MF_ALS_UNROLL8 2
#endif
#ifndef MF_ALS_UNROLL4
#define MF_ALS_UNROLL4 4
#endif
constexpr int kUnroll8 = MF_ALS_UNROLL8, kUnroll4 = MF_ALS_UNROLL4;  // rows of the Gram loop per unrolled iteration

// Work list of a half-step: one item per segment, or several for a long one — a segment with more than `split` entries
// is cut into parts of equal length (a multiple of the batch), each part accumulates its share of M on its own CTA and
// leaves it in global memory; the CTA that finishes a segment's last part adds the parts IN PART ORDER (a fixed tree,
// whoever arrives last) and factors the sum.  Without it the longest item column of the Netflix shape (180 K ratings)
// keeps one CTA busy for ~15 ms: nothing on one GPU, the whole half-step on eight.  Items are sorted longest-first.
struct AlsItem {
    uint32_t seg;     // local segment
    uint32_t part;    // part of the segment this item covers
    uint32_t nparts;  // 1: the whole segment
    uint32_t slot;    // split segments: index of part 0 in the partial-tile array (also the segment's arrival counter)
};

// 4/8/16-byte asynchronous global->shared copies (LDGSTS)
template <int BYTES>
__device__ __forceinline__ void cp_async(void* smem_dst, const void* gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    if (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
    else if (BYTES == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}

// position (in floats) of matrix index i inside a staged row / a column of L
template <int TS>
__device__ __forceinline__ int pos_of(int i, int nb) {
    const int ch = i >> 2;
    return 4 * (TS == 8 ? (ch >> 1) + (ch & 1) * nb : ch) + (i & 3);
}

// the TS entries of tile t out of one staged row / one column of L
template <int TS>
__device__ __forceinline__ void load_tile_vec(const float* __restrict__ row, int t, int nb, float (&v)[TS]) {
    const float4 lo = *reinterpret_cast<const float4*>(row + 4 * t);
    v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w;
    if (TS == 8) {
        const float4 hi = *reinterpret_cast<const float4*>(row + 4 * (nb + t));
        v[TS - 4] = hi.x; v[TS - 3] = hi.y; v[TS - 2] = hi.z; v[TS - 1] = hi.w;
    }
}

// acc += sign * a b^T
template <int TS, bool NEG>
__device__ __forceinline__ void rank1(float (&acc)[TS][TS], const float (&a)[TS], const float (&b)[TS]) {
#pragma unroll
    for (int i = 0; i < TS; ++i)
#pragma unroll
        for (int j = 0; j < TS; ++j) acc[i][j] = fmaf(NEG ? -a[i] : a[i], b[j], acc[i][j]);
}

// TS = tile edge; VW = floats per copy (4 when k % 4 == 0, 2 when k is even, else 1): every factor row starts VW-aligned.
// nb = tiles per matrix edge (nb * TS >= k + 1); ks = split-K groups; threads 0 .. ks*ntiles-1 work on the Gram matrix.
// MAXREG: register cap (sets how many CTAs fit an SM: 112 -> 3 x 192 or 6 x 96 threads, 168 -> 384 threads, 64 -> 2048).
// Staging: three buffers of kBatch rows, one barrier per batch — while batch b is accumulated, batch b+1 is landing and
// batch b+2 is being issued into the buffer batch b-1 just left.
// Shared memory is one region used in turn as staging buffers (Gram loop), split-K scratch and L (factorisation).
template <int TS, int VW, int MAXREG>
__global__ void __maxnreg__(MAXREG)
k_als_tile(int64_t nitems, const AlsItem* __restrict__ items, unsigned* __restrict__ queue, float* partial, unsigned* counters,
           const uint32_t* __restrict__ ptr, const uint32_t* __restrict__ idx, const float* __restrict__ val,
           const float* __restrict__ Y, float* __restrict__ X, int k, int nb, int ks, int region_floats, float lambda) {
    extern __shared__ __align__(16) float sm[];
    const int kp = nb * TS;
    const int ntiles = nb * (nb + 1) / 2;
    float* Ys = sm;                               // [kStages][kBatch][kp]  staged rows [factor row | rating | 0...]
    float* Lm = sm;                               // [kp][kp]               Lm[j*kp + pos(i)] = L[i][j]
    float* scratch = sm;                          // [(ks-1)][TS*TS][ntiles] split-K partial tiles
    float* Ld = sm + region_floats;               // [TS][TS]               diagonal tile being applied (natural order)
    float* dv = Ld + TS * TS;                     // [kp]                   1 / L[i][i]  (0 for i >= k)
    float* xs = dv + kp;                          // [kp]                   solution at positions pos(i)
    uint32_t* sidx = reinterpret_cast<uint32_t*>(xs + kp);  // [kStages][kBatch] row ids of the batches to be fetched

    const int tid = threadIdx.x;
    const int TPS = blockDim.x;
    const int g = tid / ntiles;                   // split-K group
    const int q = tid - g * ntiles;               // tile id, column-major over the lower triangle
    const bool active = g < ks;
    int I, J;
    {
        int off = 0;
        J = 0;
        while (q >= off + (nb - J)) { off += nb - J; ++J; }
        I = J + (q - off);
    }
    const bool owner = tid < ntiles;              // group 0 keeps the tile for the factorisation
    const int posk = pos_of<TS>(k, nb);           // where the rating sits in a staged row / y = L[k][.] in a column
    const int cpr = k / VW;                       // copies per factor row
    // staged rows are ldy floats apart: ldy = 4 mod 8, so that the 16-byte copies of 8 neighbouring lanes (one row per
    // lane) fall into 8 different bank groups
    const int ldy = kp + ((kp & 7) == 0 ? 4 : 0);
    const int lane = tid & 31, wv = tid >> 5, nwarp = TPS >> 5;

    // The queue runs one item ahead: while an item is processed, thread 0 already holds the ticket, the descriptor
    // and the extent of the next one (loads in flight), and leaves them in s_desc[parity] at the end.
    __shared__ uint32_t s_desc[2][6];  // {segment or 0xffffffff, lo, hi, part, nparts, slot}
    __shared__ unsigned s_last;
    uint32_t nx[6] = {0xffffffffu, 0, 0, 0, 1, 0};
    auto fetch_next = [&]() {
        const unsigned t = atomicAdd(queue, 1u);
        nx[0] = 0xffffffffu; nx[1] = 0; nx[2] = 0; nx[3] = 0; nx[4] = 1; nx[5] = 0;
        if (t < nitems) {
            const uint4 it = __ldg(reinterpret_cast<const uint4*>(items) + t);
            const uint32_t lo = __ldg(ptr + it.x), hi = __ldg(ptr + it.x + 1);
            const uint32_t plen = ((hi - lo + it.z - 1) / it.z + kBatch - 1) / kBatch * kBatch;  // part length (als_prepare's rule)
            const uint32_t plo = min(hi, lo + it.y * plen);
            nx[0] = it.x; nx[1] = plo; nx[2] = min(hi, plo + plen); nx[3] = it.y; nx[4] = it.z; nx[5] = it.w;
        }
    };
    auto publish_next = [&](int slot) {
#pragma unroll
        for (int i = 0; i < 6; ++i) s_desc[slot][i] = nx[i];
    };
    if (tid == 0) {
        fetch_next();
        publish_next(0);
    }
    for (int par = 0;; par ^= 1) {
        __syncthreads();  // the previous item is finished; s_desc[par] is in place
        const uint32_t seg = s_desc[par][0];
        if (seg == 0xffffffffu) break;
        const uint32_t lo = s_desc[par][1], hi = s_desc[par][2];
        const uint32_t part = s_desc[par][3], nparts = s_desc[par][4], slot = s_desc[par][5];
        if (tid == 0) fetch_next();
        float* x = X + (int64_t)seg * k;
        if (hi == lo && nparts == 1) {
            for (int c = tid; c < k; c += TPS) x[c] = 0.0f;
            if (tid == 0) publish_next(par ^ 1);
            continue;
        }
        float acc[TS][TS];
#pragma unroll
        for (int i = 0; i < TS; ++i)
#pragma unroll
            for (int j = 0; j < TS; ++j) acc[i][j] = 0.0f;

        const int nbatch = (int)((hi - lo + kBatch - 1) / kBatch);
        // the previous segment left its L in the staging region: the pad columns behind the rating have to be zero
        // again (the copies never touch them)
        for (int r = tid; r < kStages * kBatch; r += TPS)
            for (int c = k + 1; c < kp; ++c) Ys[r * ldy + pos_of<TS>(c, nb)] = 0.0f;
        // asynchronous gather of batch `b` (factor rows + ratings) into buffer `buf`, row ids from sidx: lane = row, the
        // warps share the 16-byte pieces of the row (no index arithmetic beyond one multiply per row)
        auto issue_rows = [&](int b, int buf) {
            const uint32_t base = lo + (uint32_t)b * kBatch;
            const int nrow = (int)min((uint32_t)kBatch, hi - base);
            if (lane < nrow) {
                const float* src = Y + (size_t)sidx[buf * kBatch + lane] * k;
                float* dst = Ys + (buf * kBatch + lane) * ldy;
                for (int c = wv; c < cpr; c += nwarp) cp_async<VW * 4>(dst + pos_of<TS>(c * VW, nb), src + c * VW);
                if (wv == nwarp - 1) cp_async<4>(dst + posk, val + base + lane);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        // prologue: ids of batches 0..2, rows of batches 0 and 1
        if (tid < kBatch) {
#pragma unroll
            for (int j = 0; j < kStages; ++j) {
                const uint32_t e = lo + j * kBatch + tid;
                if (e < hi) sidx[j * kBatch + tid] = __ldg(idx + e);
            }
        }
        __syncthreads();
        issue_rows(0, 0);
        if (nbatch > 1) issue_rows(1, 1); else asm volatile("cp.async.commit_group;" ::: "memory");

        int buf = 0;  // b % kStages
        for (int b = 0; b < nbatch; ++b) {
            const int nrow = (int)min((uint32_t)kBatch, hi - (lo + (uint32_t)b * kBatch));
            // ids of batch b+3 travel in a register across the barrier
            uint32_t nid = 0;
            const uint32_t e3 = lo + (uint32_t)(b + kStages) * kBatch + tid;
            const bool has3 = tid < kBatch && e3 < hi;
            if (has3) nid = __ldg(idx + e3);
            asm volatile("cp.async.wait_group 1;" ::: "memory");  // batch b has landed (batch b+1 may be in flight)
            __syncthreads();                                      // ... for everybody; everybody is done with batch b-1
            const int buf2 = buf >= 1 ? buf - 1 : kStages - 1;    // (b + 2) % kStages: the buffer batch b-1 left
            if (b + 2 < nbatch) issue_rows(b + 2, buf2); else asm volatile("cp.async.commit_group;" ::: "memory");
            const float* Yb = Ys + buf * kBatch * ldy;
            if (active) {
#pragma unroll(TS == 8 ? kUnroll8 : kUnroll4)
                for (int r = g; r < nrow; r += ks) {
                    float a[TS], c[TS];
                    load_tile_vec<TS>(Yb + r * ldy, I, nb, a);
                    load_tile_vec<TS>(Yb + r * ldy, J, nb, c);
                    rank1<TS, false>(acc, a, c);
                }
            }
            // ids of batch b+3 (loaded before the barrier, used only now: their latency hides behind the Gram loop); the
            // ids of batch b that sat here were consumed two iterations ago
            if (has3) sidx[buf * kBatch + tid] = nid;
            buf = buf + 1 == kStages ? 0 : buf + 1;
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();  // everybody is done with the staging buffers: the region becomes scratch / L

        // split-K: groups 1.. hand their partial tiles to group 0, added in group order
        if (ks > 1) {
            if (active && g > 0) {
                float* dst = scratch + (size_t)(g - 1) * TS * TS * ntiles + q;
#pragma unroll
                for (int i = 0; i < TS; ++i)
#pragma unroll
                    for (int j = 0; j < TS; ++j) dst[(i * TS + j) * ntiles] = acc[i][j];
            }
            __syncthreads();
            if (owner) {
                for (int gg = 1; gg < ks; ++gg) {
                    const float* src = scratch + (size_t)(gg - 1) * TS * TS * ntiles + q;
#pragma unroll
                    for (int i = 0; i < TS; ++i)
#pragma unroll
                        for (int j = 0; j < TS; ++j) acc[i][j] += src[(i * TS + j) * ntiles];
                }
            }
            __syncthreads();  // the scratch may reach into Lm, which the factorisation is about to write
        }
        if (nparts > 1) {
            // this CTA holds one part of the segment's M: leave it in global memory; the CTA that completes the segment
            // adds all parts in part order (its own included: the tree does not depend on who arrives last)
            if (owner) {
                float* dst = partial + (size_t)(slot + part) * TS * TS * ntiles + q;
#pragma unroll
                for (int i = 0; i < TS; ++i)
#pragma unroll
                    for (int j = 0; j < TS; ++j) dst[(i * TS + j) * ntiles] = acc[i][j];
            }
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                const unsigned arrived = atomicAdd(counters + slot, 1u);
                s_last = arrived == nparts - 1;
                if (s_last) counters[slot] = 0u;  // ready for the next half-step
            }
            __syncthreads();
            if (!s_last) {
                if (tid == 0) publish_next(par ^ 1);
                continue;
            }
            __threadfence();
            if (owner) {
#pragma unroll
                for (int i = 0; i < TS; ++i)
#pragma unroll
                    for (int j = 0; j < TS; ++j) acc[i][j] = 0.0f;
                for (uint32_t pp = 0; pp < nparts; ++pp) {
                    const float* src = partial + (size_t)(slot + pp) * TS * TS * ntiles + q;
#pragma unroll
                    for (int i = 0; i < TS; ++i)
#pragma unroll
                        for (int j = 0; j < TS; ++j) acc[i][j] += __ldcg(src + (i * TS + j) * ntiles);
                }
            }
        }
        if (owner && I == J) {
#pragma unroll
            for (int i = 0; i < TS; ++i)
                if (TS * I + i < k) acc[i][i] += lambda;
        }

        // ---- blocked right-looking Cholesky of M on the register tiles (columns >= k are switched off: dv = 0)
        for (int Jb = 0; Jb < nb; ++Jb) {
            const int jlo = Jb * TS;
            if (owner && J == Jb && I == Jb) {
                // diagonal tile: unblocked factorisation in registers
#pragma unroll
                for (int c = 0; c < TS; ++c) {
                    const float d = acc[c][c];
                    float inv = rsqrtf(d);
                    inv = inv * (1.5f - 0.5f * d * inv * inv);  // one Newton step: full FP32 accuracy
                    if (jlo + c >= k) inv = 0.0f;
                    dv[jlo + c] = inv;
                    acc[c][c] = d * inv;
#pragma unroll
                    for (int r = c + 1; r < TS; ++r) acc[r][c] *= inv;
#pragma unroll
                    for (int r = c + 1; r < TS; ++r)
#pragma unroll
                        for (int c2 = c + 1; c2 <= r; ++c2) acc[r][c2] = fmaf(-acc[r][c], acc[c2][c], acc[r][c2]);
                }
#pragma unroll
                for (int c = 0; c < TS; ++c) {
#pragma unroll
                    for (int r = 0; r < TS; ++r) Ld[r * TS + c] = r >= c ? acc[r][c] : 0.0f;
                    float* col = Lm + (jlo + c) * kp;
                    *reinterpret_cast<float4*>(col + 4 * Jb) =
                        make_float4(c <= 0 ? acc[0][c] : 0.0f, c <= 1 ? acc[1][c] : 0.0f, c <= 2 ? acc[2][c] : 0.0f, c <= 3 ? acc[3][c] : 0.0f);
                    if (TS == 8)
                        *reinterpret_cast<float4*>(col + 4 * (nb + Jb)) =
                            make_float4(c <= TS - 4 ? acc[TS - 4][c] : 0.0f, c <= TS - 3 ? acc[TS - 3][c] : 0.0f,
                                        c <= TS - 2 ? acc[TS - 2][c] : 0.0f, acc[TS - 1][c]);
                }
            }
            __syncthreads();
            if (owner && J == Jb && I > Jb) {
                // panel tile: X L_JJ^T = A_IJ, row by row
#pragma unroll
                for (int c = 0; c < TS; ++c) {
                    const float dc = dv[jlo + c];
#pragma unroll
                    for (int r = 0; r < TS; ++r) {
                        float v = acc[r][c];
#pragma unroll
                        for (int c2 = 0; c2 < c; ++c2) v = fmaf(-acc[r][c2], Ld[c * TS + c2], v);
                        acc[r][c] = v * dc;
                    }
                    float* col = Lm + (jlo + c) * kp;
                    *reinterpret_cast<float4*>(col + 4 * I) = make_float4(acc[0][c], acc[1][c], acc[2][c], acc[3][c]);
                    if (TS == 8)
                        *reinterpret_cast<float4*>(col + 4 * (nb + I)) =
                            make_float4(acc[TS - 4][c], acc[TS - 3][c], acc[TS - 2][c], acc[TS - 1][c]);
                }
            }
            __syncthreads();
            if (owner && J > Jb) {
                // trailing tile: A_IJ -= L_I,Jb L_J,Jb^T — the Gram loop with a minus sign over the TS fresh columns of L
#pragma unroll 2
                for (int c = 0; c < TS; ++c) {
                    float a[TS], b[TS];
                    load_tile_vec<TS>(Lm + (jlo + c) * kp, I, nb, a);
                    load_tile_vec<TS>(Lm + (jlo + c) * kp, J, nb, b);
                    rank1<TS, true>(acc, a, b);
                }
            }
        }
        __syncthreads();

        // ---- L^T x = y by warp 0, last block first; y_j = L[k][j] sits at position posk of column j
        if (tid < 32) {
            const int nch = kp >> 2;
            for (int Jb = nb - 1; Jb >= 0; --Jb) {
                const int jlo = Jb * TS;
                float p[TS];
#pragma unroll
                for (int c = 0; c < TS; ++c) p[c] = 0.0f;
                for (int pc = tid; pc < nch; pc += 32) {
                    const int nat = TS == 8 ? (pc < nb ? 2 * pc : 2 * (pc - nb) + 1) : pc;  // natural chunk at this position
                    if (4 * nat >= jlo + TS) {                                                // rows already solved
                        const float4 xv = *reinterpret_cast<const float4*>(xs + 4 * pc);
#pragma unroll
                        for (int c = 0; c < TS; ++c) {
                            const float4 lv = *reinterpret_cast<const float4*>(Lm + (jlo + c) * kp + 4 * pc);
                            p[c] = fmaf(lv.x, xv.x, fmaf(lv.y, xv.y, fmaf(lv.z, xv.z, fmaf(lv.w, xv.w, p[c]))));
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < TS; ++c) p[c] = warp_sum(p[c]);
                float xb[TS];
#pragma unroll
                for (int c = TS - 1; c >= 0; --c) {
                    const float* col = Lm + (jlo + c) * kp;
                    float v = col[posk] - p[c];
#pragma unroll
                    for (int c2 = c + 1; c2 < TS; ++c2) v = fmaf(-col[pos_of<TS>(jlo + c2, nb)], xb[c2], v);
                    xb[c] = (jlo + c < k) ? v * dv[jlo + c] : 0.0f;
                }
                if (tid == 0) {
#pragma unroll
                    for (int c = 0; c < TS; ++c) xs[pos_of<TS>(jlo + c, nb)] = xb[c];
                }
                __syncwarp();
            }
            for (int c = tid; c < k; c += 32) x[c] = xs[pos_of<TS>(c, nb)];
        }
        if (tid == 0) publish_next(par ^ 1);
    }
}

struct AlsGeometry {
    int TS, nb, kp, ntiles, ks, tps, region_floats;
    size_t smem;
};

AlsGeometry als_geometry(int k) {
    AlsGeometry G;
    G.TS = (k + 1 <= 24) ? 4 : 8;
    G.nb = (k + 1 + G.TS - 1) / G.TS;
    G.kp = G.nb * G.TS;
    G.ntiles = G.nb * (G.nb + 1) / 2;
    // split-K groups: small matrices need several groups to fill a CTA; from ~64 tiles on one group per CTA is best
    // (measured: k = 100 -> ks = 1, k = 40 -> ks = 4, k = 10 -> ks = 8)
    if (G.TS == 4) G.ks = std::max(1, std::min(8, 64 / G.ntiles));
    else G.ks = G.ntiles >= 64 ? 1 : std::max(1, std::min(8, 96 / G.ntiles));
    if (const char* e = getenv("MF_ALS_KS")) {  // tuning knob: split-K groups per CTA
        const int v = atoi(e);
        if (v >= 1 && v <= 8 && G.ntiles * v <= (G.TS == 8 ? 384 : 64)) G.ks = v;
    }
    G.tps = (G.ntiles * G.ks + 31) / 32 * 32;
    const int ldy = G.kp + (G.kp % 8 == 0 ? 4 : 0);
    const size_t stage = (size_t)kStages * kBatch * ldy, lmat = (size_t)G.kp * G.kp;
    const size_t scratch = (size_t)(G.ks - 1) * G.TS * G.TS * G.ntiles;
    G.region_floats = (int)((std::max(std::max(stage, lmat), scratch) + 3) / 4 * 4);
    G.smem = sizeof(float) * ((size_t)G.region_floats + G.TS * G.TS + 2 * (size_t)G.kp) + sizeof(uint32_t) * kStages * kBatch;
    return G;
}

template <int TS, int VW, int MAXREG>
int launch_als(const AlsGeometry& G, const Side& s, const float* Y, float* X,
               int k, float lambda, int sm_count, cudaStream_t st) {
    // once per half-step (two host calls): shared-memory limit for this k, and all of the SM's configurable memory as
    // shared memory — the CTA count per SM is what hides the serial phases
    if (G.smem > 48 * 1024)
        MF_CUDA(cudaFuncSetAttribute(k_als_tile<TS, VW, MAXREG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G.smem));
    MF_CUDA(cudaFuncSetAttribute(k_als_tile<TS, VW, MAXREG>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int per_sm = 1;
    MF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_als_tile<TS, VW, MAXREG>, G.tps, G.smem));
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)sm_count * per_sm;
    if (grid > s.als_nitems) grid = s.als_nitems;
    k_als_tile<TS, VW, MAXREG><<<(unsigned)grid, G.tps, G.smem, st>>>(s.als_nitems, reinterpret_cast<const AlsItem*>(s.als_items),
                                                                        s.als_queue, s.als_partial, s.als_counters, s.ptr, s.idx,
                                                                        s.val, Y, X, k, G.nb, G.ks, G.region_floats, lambda);
    MF_CUDA(cudaGetLastError());
    return MF_OK;
}

template <int TS, int VW>
int launch_als_vw(const AlsGeometry& G, const Side& s, const float* Y,
                  float* X, int k, float lambda, int sm_count, cudaStream_t st) {
    // register classes: 64 (4 x 4 tiles: any number of 32/64-thread CTAs), 112 (8 x 8 tiles: 3 x 192 or 6 x 96 threads per
    // SM), 168 (8 x 8 tiles, up to 384 threads)
    const char* fr = getenv("MF_ALS_REGS");  // tuning knob: 112 or 168
    const int force_regs = fr ? atoi(fr) : 0;
    if constexpr (TS == 4) return launch_als<TS, VW, MF_ALS_REG4>(G, s, Y, X, k, lambda, sm_count, st);
    else {
        const bool small = force_regs ? force_regs < 168 : G.tps <= 192;
        if (small && G.tps <= 192) return launch_als<TS, VW, MF_ALS_REG8>(G, s, Y, X, k, lambda, sm_count, st);
        return launch_als<TS, VW, 168>(G, s, Y, X, k, lambda, sm_count, st);
    }
}

template <int TS>
int launch_als_ts(const AlsGeometry& G, const Side& s, const float* Y,
                  float* X, int k, float lambda, int sm_count, cudaStream_t st) {
    if (k % 4 == 0) return launch_als_vw<TS, 4>(G, s, Y, X, k, lambda, sm_count, st);
    if (k % 2 == 0) return launch_als_vw<TS, 2>(G, s, Y, X, k, lambda, sm_count, st);
    return launch_als_vw<TS, 1>(G, s, Y, X, k, lambda, sm_count, st);
}

}  // namespace

// The work list of a half-step from the (host) pointer array: one item per segment, or `nparts` items for a segment with
// more than `split` entries; parts have equal length plen = roundup32(ceil(deg / nparts)) — the rule the kernel re-derives
// from (deg, nparts) — and the list is sorted longest-first (stable).  *slots = number of partial-tile slots needed.
void als_plan(const uint32_t* hptr, int64_t nseg, uint32_t split, std::vector<AlsItem>& sorted, uint32_t* slots_out) {
    std::vector<AlsItem> items;
    std::vector<uint32_t> len;
    items.reserve((size_t)nseg);
    len.reserve((size_t)nseg);
    uint32_t slots = 0;
    for (int64_t sg = 0; sg < nseg; ++sg) {
        const uint32_t deg = hptr[sg + 1] - hptr[sg];
        uint32_t nparts = deg > split ? (deg + split - 1) / split : 1;
        uint32_t plen = nparts > 1 ? ((deg + nparts - 1) / nparts + kBatch - 1) / kBatch * kBatch : deg;
        if (nparts > 1) nparts = (deg + plen - 1) / plen;  // rounding the part length up may save a part
        if (nparts <= 1) { nparts = 1; plen = deg; }
        // the kernel derives the part length from (deg, nparts) by the same rule: make sure both agree
        if (nparts > 1 && ((deg + nparts - 1) / nparts + kBatch - 1) / kBatch * kBatch != plen) { nparts = 1; plen = deg; }
        for (uint32_t p = 0; p < nparts; ++p) {
            items.push_back({(uint32_t)sg, p, nparts, nparts > 1 ? slots : 0u});
            len.push_back(nparts > 1 ? std::min(plen, deg - p * plen) : deg);
        }
        if (nparts > 1) slots += nparts;
    }
    std::vector<uint32_t> perm(items.size());
    for (size_t i = 0; i < perm.size(); ++i) perm[i] = (uint32_t)i;
    std::stable_sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return len[a] > len[b]; });
    sorted.resize(items.size());
    for (size_t i = 0; i < perm.size(); ++i) sorted[i] = items[perm[i]];
    *slots_out = slots;
}

// the work list of a side (items sorted longest-first, long segments split), built on first use and kept for the life
// of the session; host-side: one download of the pointer array, one sort
int als_prepare(Side& s, const AlsGeometry& G, cudaStream_t st) {
    if (s.als_items || s.nseg <= 0) return MF_OK;
    std::vector<uint32_t> hptr((size_t)s.nseg + 1);
    MF_CUDA(cudaMemcpyAsync(hptr.data(), s.ptr, sizeof(uint32_t) * hptr.size(), cudaMemcpyDeviceToHost, st));
    MF_CUDA(cudaStreamSynchronize(st));
    uint32_t split = 8192;  // entries per part (~0.7 ms of Gram work at k = 100); measured flat from 4 K to 64 K on one GPU at
                            // k = 40 / 100, 6 % better than 16 K at k = 10, and shorter tails when the shard is small
    if (const char* e = getenv("MF_ALS_SPLIT")) {  // tuning / test knob
        const long v = atol(e);
        if (v >= kBatch) split = (uint32_t)v;
    }
    std::vector<AlsItem> sorted;
    uint32_t slots = 0;
    als_plan(hptr.data(), s.nseg, split, sorted, &slots);
    s.als_nitems = (int64_t)sorted.size();
    AlsItem* d_items = nullptr;
    MF_TRY(dev_alloc(&d_items, sorted.size()));
    s.als_items = d_items;
    MF_TRY(dev_alloc(&s.als_queue, 1));
    MF_TRY(dev_alloc(&s.als_counters, (size_t)slots + 1));
    MF_TRY(dev_alloc(&s.als_partial, (size_t)slots * G.TS * G.TS * G.ntiles + 1));
    MF_CUDA(cudaMemcpyAsync(d_items, sorted.data(), sizeof(AlsItem) * sorted.size(), cudaMemcpyHostToDevice, st));
    MF_CUDA(cudaMemsetAsync(s.als_counters, 0, sizeof(unsigned) * ((size_t)slots + 1), st));
    MF_CUDA(cudaStreamSynchronize(st));  // `sorted` goes out of scope
    return MF_OK;
}

int als_half_step(Side& s, const float* Y, float* X, int k, float lambda, int sm_count, cudaStream_t st) {
    if (s.nseg <= 0) return MF_OK;
    const AlsGeometry G = als_geometry(k);
    if (G.smem > 227 * 1024 || G.tps > (G.TS == 8 ? 384 : 64)) {
        set_error("ALS: k=%d needs %zu bytes of shared memory per CTA (limit 227 KB)", k, G.smem);
        return MF_ERR_UNSUPPORTED;
    }
    MF_TRY(als_prepare(s, G, st));
    MF_CUDA(cudaMemsetAsync(s.als_queue, 0, sizeof(unsigned), st));
    if (G.TS == 8) return launch_als_ts<8>(G, s, Y, X, k, lambda, sm_count, st);
    return launch_als_ts<4>(G, s, Y, X, k, lambda, sm_count, st);
}

// host-only planner behind the C-ABI (include/mf_abi.h: mf_als_plan)
int als_plan_host(const uint32_t* ptr, int64_t nseg, uint32_t split, uint32_t* items4, int64_t* n_items, uint32_t* n_slots) {
    if (split < (uint32_t)kBatch) split = (uint32_t)kBatch;
    std::vector<AlsItem> sorted;
    uint32_t slots = 0;
    als_plan(ptr, nseg, split, sorted, &slots);
    if (n_items) *n_items = (int64_t)sorted.size();
    if (n_slots) *n_slots = slots;
    if (items4) memcpy(items4, sorted.data(), sizeof(AlsItem) * sorted.size());
    return MF_OK;
}

}  // namespace mf
