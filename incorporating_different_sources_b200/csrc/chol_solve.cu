// Batched blocked Cholesky factorisation + triangular solves + posterior scalars.
//
// Replaces, per window, the reference's explicit inverse and mat-vec
//   w1 = inv(S1) (c S0 w0 + t)                      portfolio_calculations.py:485-489
//   nu = (n1 + N + 2) w1 / (n1 - w1' S1 w1)         :572-575      weights = (1/gamma) nu   :836
//   nu_J = inv(T - t t'/n) t                        :600-606      weights = (1/gamma) nu_J :849
// by S = L L', L z = b, L' w = z (within 1.5e-12 of inv()*b on well-posed inputs) and
// v1 = w1' S1 w1 = z'z.
//
// One CTA (4 warps) per window, 6 CTAs per SM (measured: 4 -> 29.7 ms, 6 -> 26.4 ms, 8 -> 26.5 ms per step), left-looking blocked factorisation with 32-column
// panels, in place in the [rows][ldS] device layout produced by the Gram kernel (lower triangle):
//   U  panel update   C = S[j0:, j0:j0+32] - L[j0:, :j0] L[j0:j0+32, :j0]'   on the FP64 tensor cores.
//                     The already factored columns are streamed through shared memory by TMA
//                     (2-D tensor map over the whole workspace, 32-row x 16-column boxes,
//                     SWIZZLE_128B, 3-stage mbarrier ring) in slabs of 8*TPW m-tiles; each warp owns TPW
//                     8-row m-tiles of the slab and all four 8-column n-tiles of the panel.
//   F  diagonal block 32x32 right-looking Cholesky AND its triangular inverse in ONE 32-step loop by the
//                     whole CTA in shared memory (one barrier per step; a single warp doing this alone is
//                     latency bound and took 2/3 of the kernel).  The diagonal block lives in the first
//                     slab, so F runs right after that slab's update.
//   T  panel solve    L[j0+32:, j0:j0+32] = C * inv(L_d)' on the tensor cores, FUSED into the epilogue of
//                     the panel update: the accumulator fragment of a DMMA is, column for column, a valid
//                     A fragment (lane (g,tig) holds columns 8nt+2tig+h, which become the k slots of step
//                     (nt,h)), so C never goes to memory: the accumulators start as -S, collect +L L',
//                     are multiplied by inv(L_d)' from registers, and the finished factor rows are stored
//                     once.  (The first version wrote C and re-read it in a separate T phase: 2 MB of DRAM
//                     traffic per window and 15% of the kernel.)
// The factor is written with generic stores and re-read by TMA in later panels, so every panel ends
// with fence.proxy.async + a block barrier.  The right-hand side rides along as one extra row
// (row Nr = roundup(N,32)) of the matrix, so the forward substitution L z = b is a by-product of
// the factorisation.  The back substitution then walks the panels in reverse with coalesced row
// reads and a warp-level triangular solve.
//
// k-permutation: inside a 16-column chunk, DMMA step q contracts columns {2q, 2q+1, 8+2q, 9+2q}
// over the four tig lanes.  With the 128-byte swizzle (16-byte unit index XOR row&7) the 16 lanes of
// a half warp then hit 16 distinct 8-byte words of the 128-byte bank line: every LDS.64 fragment
// load is conflict free.  A and B fragments use the same assignment, so the contraction is exact.
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace bp {

constexpr int NB = 32;

#ifndef CH_NWARPS
#define CH_NWARPS 4
#endif
constexpr int CH_WARPS = CH_NWARPS;                 // warps per window (CTA)
#ifndef CH_CTAS_PER_SM
#define CH_CTAS_PER_SM 6
#endif
constexpr int CH_OCC = CH_CTAS_PER_SM;              // resident CTAs (windows) per SM
constexpr int CH_THREADS = CH_WARPS * 32;

constexpr int LDQ = 34;    // row stride of the packed diagonal-block array

// 32x32 Cholesky AND triangular inverse of the diagonal block by the WHOLE CTA, one barrier per step.
// In : C (lower triangle) at P[r*LDQ + c], c <= r.
// Out: L in the same place; inv(L) TRANSPOSED and shifted by one column in the strictly upper part,
//      inv(L)[n][k] at P[k*LDQ + n + 1] (k <= n) -- the layout in which the B fragments of the fused panel
//      solve are bank-conflict free (8-byte bank (4*tig + g + const) mod 16); 1/diag in invd.
// Returns (to every thread) the 1-based index of the first non-positive pivot, or 0.
//
// Step k, Cholesky part (thread = row i, EPT consecutive columns): reads the UNSCALED column k and the
// pivot, derives its own l_i and l_j (rsqrt recomputed by every thread instead of being broadcast through
// a second barrier) and updates its trailing elements k < j <= i.  The scaled column k is written one step
// later, after the barrier, when nobody reads column k any more.
// Step k, inverse part (thread = inverse row iy = lane, EPT inverse columns of its warp): with
// Y[i][j] = inv(L)[i][j] * L[i][i] (row scaling deferred to the end), Y[i][k] = -m_ik and
// Y[i][j] -= m_ik Y[k][j] for j < k < i, m_ik = A[i][k] / pivot.  Reads column k+1, writes columns > k+1.
__device__ __forceinline__ int potrf_trtri_block(double* P, double* invd, int tid) {
    constexpr int TPR = CH_THREADS / NB;   // threads per row of the block
    constexpr int EPT = NB / TPR;          // consecutive elements owned by a thread
    static_assert(TPR == CH_WARPS, "inverse part: one warp per group of EPT inverse columns");
    const int i = tid / TPR;               // Cholesky part: row owned by this thread
    const int jb = (tid % TPR) * EPT;      //                its EPT consecutive columns
    const int iy = tid & 31;               // inverse part: inverse row owned by this thread
    const int jy = (tid >> 5) * EPT;       //               its EPT inverse columns (warp-uniform)
    int fail = 0;
    double lprev = 0.0;
    for (int k = 0; k < NB; ++k) {
        const double piv = P[k * LDQ + k];
        if (!(piv > 0.0) && fail == 0) fail = k + 1;
        const double inv = rsqrt(piv);
        const double li = i > k ? P[i * LDQ + k] * inv : (i == k ? piv * inv : 0.0);
        const double my = iy > k ? P[iy * LDQ + k] * inv * inv : 0.0;
        if (k > 0 && k - 1 >= jb && k - 1 < jb + EPT && i >= k - 1) P[i * LDQ + k - 1] = lprev;   // deferred column k-1
        if (k >= jb && k < jb + EPT) lprev = li;
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            const int j = jb + e;
            if (j > k && j <= i) P[i * LDQ + j] = fma(-li, P[j * LDQ + k] * inv, P[i * LDQ + j]);
        }
        if (iy > k) {
#pragma unroll
            for (int e = 0; e < EPT; ++e) {
                const int j = jy + e;
                if (j < k) P[j * LDQ + iy + 1] = fma(-my, P[j * LDQ + k + 1], P[j * LDQ + iy + 1]);
                else if (j == k) P[j * LDQ + iy + 1] = -my;
            }
        }
        if (tid == 0) invd[k] = inv;
        __syncthreads();
    }
    if (NB - 1 >= jb && NB - 1 < jb + EPT && i >= NB - 1) P[i * LDQ + NB - 1] = lprev;
    // row scaling of the inverse: inv(L)[i][j] = Y[i][j] / L[i][i], inv(L)[i][i] = 1 / L[i][i]
    {
        const double d = invd[iy];
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            const int j = jy + e;
            if (j < iy) P[j * LDQ + iy + 1] *= d;
            else if (j == iy) P[j * LDQ + iy + 1] = d;
        }
    }
    __syncthreads();
    return fail;
}

#ifndef CH_TPW
#define CH_TPW 2                                     // m-tiles per warp and slab
#endif
#ifndef CH_NSTAGES
#define CH_NSTAGES 2
#endif
constexpr int TPW = CH_TPW;
constexpr int SLAB_TILES = CH_WARPS * TPW;           // m-tiles per slab (one pass of the panel update)
constexpr int ABOXES = SLAB_TILES / 4;               // 32-row TMA boxes per slab
constexpr int CH_STAGES = CH_NSTAGES;
constexpr int BOX_BYTES = 32 * 16 * 8;               // one TMA box: 32 rows x 16 columns of doubles
constexpr int CH_STAGE_BYTES = (ABOXES + 1) * BOX_BYTES;   // A slab + 1 B box
constexpr int CH_TMA_SMEM = CH_STAGES * CH_STAGE_BYTES + 1024;

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
// generic-proxy global writes (the factor) must be visible to later async-proxy (TMA) reads
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

__global__ void __launch_bounds__(CH_THREADS, CH_OCC)
chol_solve_kernel(const __grid_constant__ CUtensorMap smap, const SolveParams p) {
    extern __shared__ unsigned char sm_raw[];
    __shared__ uint64_t full_bar[CH_STAGES];
    __shared__ uint64_t empty_bar[CH_STAGES];     // one arrival per consumer warp
    __shared__ int fail_s;
    // 1024-byte alignment by pointer arithmetic on the shared array (keeps the shared address space)
    unsigned char* stage_mem = sm_raw + ((1024u - (smem_u32(sm_raw) & 1023u)) & 1023u);
    double* P = reinterpret_cast<double*>(stage_mem + CH_STAGES * CH_STAGE_BYTES);   // [32][LDQ]: L_d (lower) + inv(L_d)' (upper)
    double* invd = P + NB * LDQ;           // [32] reciprocal diagonal of the current block
    double* red = invd + NB;               // [CH_WARPS][32]
    double* scratch = red + CH_WARPS * NB; // [40]
    // solution vector of the back substitution: aliases the TMA stages (after the last panel they are idle)
    double* xs = reinterpret_cast<double*>(stage_mem);   // [Nr]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, tig = lane & 3;
    const int N = p.n_assets;
    const int ld = p.ldS;
    const int Nr = (N + NB - 1) / NB * NB;     // row index of the right-hand side
    const int rowsS = (int)(p.win_stride / ld);

    if (tid == 0) {
        tma_prefetch_desc(&smap);
        for (int s = 0; s < CH_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], CH_WARPS);
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    __syncthreads();
    // per-lane fragment offsets inside a [32][16] swizzled box: row r = 8*t + g, DMMA step q reads column
    // c = 2q + (tig&1) + 8*(tig>>1): 16-byte unit (q + 4*(tig>>1)) ^ g, 8-byte half tig&1
    int foff[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) foff[q] = g * 128 + (((q + 4 * (tig >> 1)) ^ g) << 4) + (tig & 1) * 8;
    uint32_t it = 0;      // chunks consumed so far (mbarrier phase bookkeeping)
    // optional phase profile (BP_CHOL_PROFILE=1): cycles of thread 0 between phase boundaries
    long long prof[6] = {0, 0, 0, 0, 0, 0};
    long long tlast = p.debug ? clock64() : 0;
#define PROF(k)                                   \
    if (p.debug && tid == 0) {                    \
        const long long tn = clock64();           \
        prof[k] += tn - tlast;                    \
        tlast = tn;                               \
    }

    for (int w = blockIdx.x; w < p.n_windows; w += gridDim.x) {
        double* S = p.S + (long long)w * p.win_stride;
        const double* rhs = p.rhs + (long long)w * p.ldv;
        const int grow0 = w * rowsS;           // first row of this window in the tensor map
        if (tid == 0) fail_s = 0;
        for (int j = tid; j < ld; j += CH_THREADS) S[(long long)Nr * ld + j] = j < N ? rhs[j] : 0.0;
        fence_proxy_async_all();
        __syncthreads();

        for (int j0 = 0; j0 < N; j0 += NB) {
            const int mt_total = (Nr + 8 - j0) / 8;        // m-tiles covering rows j0 .. Nr+7
            const int nchunks = j0 / 16;
            // ---------------- U (+F after the first slab) + fused T, one slab of SLAB_TILES m-tiles per pass
            for (int slab0 = 0; slab0 < mt_total; slab0 += SLAB_TILES) {
                const int slab_tiles = min(SLAB_TILES, mt_total - slab0);
                const int nboxes = (slab_tiles + 3) >> 2;
                const int slab_row = j0 + 8 * slab0;
                int ni = 0;                  // m-tiles this warp really has in this slab (warp-uniform)
#pragma unroll
                for (int i = 0; i < TPW; ++i) ni += (warp + CH_WARPS * i) < slab_tiles ? 1 : 0;

                auto issue = [&](int c, uint32_t seq) {
                    const int stage = seq % CH_STAGES;
                    unsigned char* dst = stage_mem + stage * CH_STAGE_BYTES;
                    // the stage is free once all consumer warps have released its previous use
                    if (seq >= CH_STAGES) mbar_wait(&empty_bar[stage], ((seq / CH_STAGES) - 1) & 1);
                    mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(nboxes + 1) * BOX_BYTES);
                    for (int bx = 0; bx < nboxes; ++bx)
                        tma_load_2d(dst + bx * BOX_BYTES, &smap, 16 * c, grow0 + slab_row + 32 * bx, &full_bar[stage]);
                    tma_load_2d(dst + ABOXES * BOX_BYTES, &smap, 16 * c, grow0 + j0, &full_bar[stage]);
                };
                if (tid == 0)
                    for (int c = 0; c < CH_STAGES && c < nchunks; ++c) issue(c, it + c);

                // accumulators start as -S (identity rows for the padding), collect +L L': acc = -C
                double acc[TPW][4][2];
#pragma unroll
                for (int i = 0; i < TPW; ++i) {
                    const int row = j0 + 8 * (slab0 + warp + CH_WARPS * i) + g;
                    const bool real = i < ni && (row < N || row == Nr);
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) {
                        const int col = j0 + 8 * nt + 2 * tig;
                        double2 sv = make_double2(0.0, 0.0);
                        if (real && col < N) sv = *reinterpret_cast<const double2*>(S + (long long)row * ld + col);
                        if (real) {
                            acc[i][nt][0] = -sv.x;
                            acc[i][nt][1] = col + 1 < N ? -sv.y : 0.0;
                        } else {
                            acc[i][nt][0] = row == col ? -1.0 : 0.0;
                            acc[i][nt][1] = row == col + 1 ? -1.0 : 0.0;
                        }
                    }
                }

                for (int c = 0; c < nchunks; ++c, ++it) {
                    const int stage = it % CH_STAGES;
                    const unsigned char* sA = stage_mem + stage * CH_STAGE_BYTES;
                    const unsigned char* sB = sA + ABOXES * BOX_BYTES;
                    mbar_wait(&full_bar[stage], (it / CH_STAGES) & 1);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        double b[4];
#pragma unroll
                        for (int nt = 0; nt < 4; ++nt)
                            b[nt] = *reinterpret_cast<const double*>(sB + nt * 1024 + foff[q]);
#pragma unroll
                        for (int i = 0; i < TPW; ++i) {
                            if (i < ni) {    // skip the m-tile slots past the end of the slab
                                const int t = warp + CH_WARPS * i;
                                const double a = *reinterpret_cast<const double*>(sA + (t >> 2) * BOX_BYTES + (t & 3) * 1024 + foff[q]);
#pragma unroll
                                for (int nt = 0; nt < 4; ++nt) dmma884(acc[i][nt][0], acc[i][nt][1], a, b[nt]);
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty_bar[stage]);   // this warp is done with the stage
                    if (tid == 0 && c + CH_STAGES < nchunks) issue(c + CH_STAGES, it + CH_STAGES);
                }

                // padding: rows N..Nr-1 and columns >= N of the workspace hold no data; whatever the update
                // accumulated there is replaced by the identity (rows) / zero (columns)
                if (slab_row + 8 * slab_tiles > N || j0 + NB > N) {
#pragma unroll
                    for (int i = 0; i < TPW; ++i) {
                        const int row = j0 + 8 * (slab0 + warp + CH_WARPS * i) + g;
                        const bool real = row < N || row == Nr;
#pragma unroll
                        for (int nt = 0; nt < 4; ++nt) {
                            const int col = j0 + 8 * nt + 2 * tig;
                            if (!real) {
                                acc[i][nt][0] = row == col ? -1.0 : 0.0;
                                acc[i][nt][1] = row == col + 1 ? -1.0 : 0.0;
                            } else {
                                if (col >= N) acc[i][nt][0] = 0.0;
                                if (col + 1 >= N) acc[i][nt][1] = 0.0;
                            }
                        }
                    }
                }

                if (slab0 == 0) {
                    // ---------------- F: the diagonal block (m-tiles 0..3 of the first slab) goes to shared memory
#pragma unroll
                    for (int i = 0; i < TPW; ++i) {
                        const int q = warp + CH_WARPS * i;
                        if (i < ni && q < NB / 8) {
                            const int lr = 8 * q + g;
#pragma unroll
                            for (int nt = 0; nt < 4; ++nt) {
                                const int lc = 8 * nt + 2 * tig;
                                if (lc <= lr) P[lr * LDQ + lc] = -acc[i][nt][0];
                                if (lc + 1 <= lr) P[lr * LDQ + lc + 1] = -acc[i][nt][1];
                            }
                        }
                    }
                    __syncthreads();
                    PROF(0)
                    const int f = potrf_trtri_block(P, invd, tid);
                    if (tid == 0 && f != 0 && fail_s == 0) fail_s = j0 + f;
                    PROF(1)
                    // write L_d back (lower part, real rows only)
                    for (int i = warp; i < NB; i += CH_WARPS) {
                        const int row = j0 + i, col = j0 + lane;
                        if (row < N && col < N && lane <= i) S[(long long)row * ld + col] = P[i * LDQ + lane];
                    }
                }
                // ---------------- T: X = C inv(L_d)' straight from the accumulators.  DMMA step (kt,h) contracts
                // the columns 8kt+2tig+h the lanes already hold; B[k][n] = inv(L_d)[n][k] = P[k*LDQ + n + 1];
                // inv(L_d)[n][k] = 0 for k > n, so n-tiles left of the k block are skipped
#pragma unroll
                for (int i = 0; i < TPW; ++i) {
                    const int q = slab0 + warp + CH_WARPS * i;     // m-tile index within the panel
                    if (i >= ni || q < NB / 8) continue;
                    const int row = j0 + 8 * q + g;
                    double y[4][2];
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) y[nt][0] = y[nt][1] = 0.0;
#pragma unroll
                    for (int kt = 0; kt < 4; ++kt)
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int k = 8 * kt + 2 * tig + h;
#pragma unroll
                            for (int nt = kt; nt < 4; ++nt) {
                                double bv = P[k * LDQ + 8 * nt + g + 1];
                                if (nt == kt && 2 * tig + h > g) bv = 0.0;
                                dmma884(y[nt][0], y[nt][1], acc[i][kt][h], bv);
                            }
                        }
                    if (row < N || row == Nr) {
#pragma unroll
                        for (int nt = 0; nt < 4; ++nt) {
                            const int col = j0 + 8 * nt + 2 * tig;
                            if (col < N)
                                *reinterpret_cast<double2*>(S + (long long)row * ld + col) = make_double2(-y[nt][0], -y[nt][1]);
                        }
                    }
                }
            }
            fence_proxy_async_all();        // the factor written above is read by TMA in the next panels
            __syncthreads();
            PROF(2)
        }

        // ---------------- z = L^-1 b sits in row Nr;  v1 = z'z = w1' S1 w1   (:574)
        double zz = 0.0;
        for (int j = tid; j < Nr; j += CH_THREADS) {
            const double z = j < N ? S[(long long)Nr * ld + j] : 0.0;
            xs[j] = z;
            zz = fma(z, z, zz);
        }
        const double v1 = block_sum(zz, scratch);

        // ---------------- back substitution  L' x = z, panels in reverse
        for (int j0 = Nr - NB; j0 >= 0; j0 -= NB) {
            double part = 0.0;
            const int col = j0 + lane;
            if (col < N) {
                int i = j0 + NB + warp;
                for (; i + 3 * CH_WARPS < N; i += 4 * CH_WARPS) {        // 4 independent loads in flight
                    const double s0 = S[(long long)i * ld + col], s1 = S[(long long)(i + CH_WARPS) * ld + col];
                    const double s2 = S[(long long)(i + 2 * CH_WARPS) * ld + col], s3 = S[(long long)(i + 3 * CH_WARPS) * ld + col];
                    part = fma(s0, xs[i], part);
                    part = fma(s1, xs[i + CH_WARPS], part);
                    part = fma(s2, xs[i + 2 * CH_WARPS], part);
                    part = fma(s3, xs[i + 3 * CH_WARPS], part);
                }
                for (; i < N; i += CH_WARPS) part = fma(S[(long long)i * ld + col], xs[i], part);
            }
            red[warp * NB + lane] = part;
            for (int i = warp; i < NB; i += CH_WARPS) {
                const int row = j0 + i;
                double v;
                if (row < N && col < N) v = lane <= i ? S[(long long)row * ld + col] : 0.0;
                else v = i == lane ? 1.0 : 0.0;
                P[i * LDQ + lane] = v;
            }
            __syncthreads();
            if (warp == 0) {
                double r = xs[col];
#pragma unroll
                for (int wv = 0; wv < CH_WARPS; ++wv) r -= red[wv * NB + lane];
                const double rd = 1.0 / P[lane * LDQ + lane];      // all 32 reciprocals in parallel
#pragma unroll
                for (int k = NB - 1; k >= 0; --k) {
                    const double xk = __shfl_sync(0xffffffffu, r * rd, k);
                    if (lane == k) r = xk;
                    if (lane < k) r = fma(-P[k * LDQ + lane], xk, r);
                }
                xs[col] = r;
            }
            __syncthreads();
        }

        PROF(3)
        // ---------------- posterior scalars and weights
        double* scal = p.scal + (long long)w * BP_S_COUNT;
        double mult = 1.0;
        if (p.mode == BP_MODE_CONJUGATE) {
            const double n1 = scal[BP_S_N1];
            mult = (n1 + (double)N + 2.0) / (n1 - v1);
        }
        for (int j = tid; j < p.ldv; j += CH_THREADS) {
            const double wv = j < N ? xs[j] : 0.0;
            const double nu = p.mode == BP_MODE_CONJUGATE ? (wv * mult) : wv;
            p.w1[(long long)w * p.ldv + j] = wv;
            p.nu[(long long)w * p.ldv + j] = nu;
            p.weights[(long long)w * p.ldv + j] = p.inv_gamma * nu;
        }
        if (tid == 0) {
            scal[BP_S_V1] = v1;
            p.status[w] = fail_s;
        }
        __syncthreads();
        PROF(4)
    }
    if (p.debug && tid == 0)
        for (int k = 0; k < 6; ++k) atomicAdd(reinterpret_cast<unsigned long long*>(p.debug) + k, (unsigned long long)prof[k]);
#undef PROF
}

size_t chol_smem_bytes(int n_assets) {
    const int Nr = (n_assets + NB - 1) / NB * NB;
    // xs aliases the stage ring: it must hold it
    if ((size_t)CH_STAGES * CH_STAGE_BYTES < sizeof(double) * (size_t)Nr) return 0;
    return (size_t)CH_TMA_SMEM + sizeof(double) * (size_t)(NB * LDQ + NB + CH_WARPS * NB + 40);
}

cudaError_t launch_chol_solve(const SolveParams& p, const CUtensorMap& smap, int sm_count, cudaStream_t st) {
    if (p.n_windows <= 0) return cudaSuccess;
    const size_t smem = chol_smem_bytes(p.n_assets);
    if (smem == 0) return cudaErrorInvalidValue;      // N too large for the aliased back-substitution vector
    cudaError_t e = cudaFuncSetAttribute(chol_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    static const int ctas_per_sm = [] {
        const char* e = getenv("BP_CHOL_CTAS_PER_SM");      // tuning knob (default: CH_OCC)
        const int v = e ? atoi(e) : CH_OCC;
        return v >= 1 && v <= CH_OCC ? v : CH_OCC;
    }();
    int grid = ctas_per_sm * sm_count;
    if (grid > p.n_windows) grid = p.n_windows;
    static const bool profile = getenv("BP_CHOL_PROFILE") != nullptr;
    if (profile) {
        SolveParams q = p;
        static long long* dbg = nullptr;
        if (!dbg) cudaMalloc(&dbg, 6 * sizeof(long long));
        cudaMemsetAsync(dbg, 0, 6 * sizeof(long long), st);
        q.debug = dbg;
        chol_solve_kernel<<<grid, CH_THREADS, smem, st>>>(smap, q);
        long long hostv[6];
        cudaMemcpyAsync(hostv, dbg, sizeof(hostv), cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        const double tot = (double)(hostv[0] + hostv[1] + hostv[2] + hostv[3] + hostv[4]);
        fprintf(stderr, "[chol profile] W=%d U(slab0)=%.1f%% F=%.1f%% U+T(rest)=%.1f%% backsub=%.1f%% out=%.1f%% cycles/window=%.0f\n", p.n_windows,
                100 * hostv[0] / tot, 100 * hostv[1] / tot, 100 * hostv[2] / tot, 100 * hostv[3] / tot, 100 * hostv[4] / tot,
                tot / p.n_windows);
        return cudaGetLastError();
    }
    chol_solve_kernel<<<grid, CH_THREADS, smem, st>>>(smap, p);
    return cudaGetLastError();
}

}  // namespace bp
