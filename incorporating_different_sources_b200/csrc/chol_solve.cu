// Batched blocked Cholesky factorisation + triangular solves + posterior scalars.
//
// Replaces, per window, the reference's explicit inverse and mat-vec
//   w1 = inv(S1) (c S0 w0 + t)                      portfolio_calculations.py:485-489
//   nu = (n1 + N + 2) w1 / (n1 - w1' S1 w1)         :572-575      weights = (1/gamma) nu   :836
//   nu_J = inv(T - t t'/n) t                        :600-606      weights = (1/gamma) nu_J :849
// by S = L L', L z = b, L' w = z (within 1.5e-12 of inv()*b on well-posed inputs) and
// v1 = w1' S1 w1 = z'z.
//
// One CTA per window, left-looking blocked factorisation with 32-column panels, in place in the
// [rows][ldS] device layout produced by the Gram kernel (lower triangle):
//   U  panel update   C = S[j0:, j0:j0+32] - L[j0:, :j0] L[j0:j0+32, :j0]'     DMMA, fragments read
//                     straight from global/L2 as 32-byte vectors (k-permuted, see below)
//   F  diagonal block 32x32 right-looking Cholesky by one warp (row in registers, warp shuffles,
//                     rsqrt pivots: the dependent FP64 chain per column is 4 operations)
//   T  panel solve    L[j0+32:, j0:j0+32] = C * inv(L_d)', one thread per row, L_d broadcast
//                     from shared memory (same FP64 rate as DMMA on B200, no triangular inverse)
// The right-hand side rides along as one extra row (row Nr = roundup(N,32)) of the matrix, so the
// forward substitution L z = b is a by-product of the factorisation.  The back substitution then
// walks the panels in reverse with coalesced row reads and a warp-level triangular solve.
//
// k-permutation: of every 16-wide k chunk a lane (g, tig) loads columns {2tig, 2tig+1, 8+2tig,
// 9+2tig} of its row (two 16-byte loads; a warp-wide load covers whole 32-byte sectors); DMMA step
// q (0..3) contracts the q-th of those four columns over the four tig lanes.  A and B fragments use
// the same assignment, so the contraction over the chunk is exact.
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace bp {

constexpr int NB = 32;
constexpr int LDP = 33;    // padded shared-memory row stride of the 32x32 blocks

struct d4 {
    double v[4];
};

// ptr addresses column k0 + 2*tig of the lane's row: the lane takes columns {2tig, 2tig+1, 8+2tig, 9+2tig}
// of the 16-wide chunk, so each of the two 16-byte loads of a warp covers whole 32-byte sectors
__device__ __forceinline__ d4 load4(const double* ptr, bool pred) {
    d4 r;
    if (pred) {
        const double2 lo = *reinterpret_cast<const double2*>(ptr);
        const double2 hi = *reinterpret_cast<const double2*>(ptr + 8);
        r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = hi.x; r.v[3] = hi.y;
    } else {
        r.v[0] = r.v[1] = r.v[2] = r.v[3] = 0.0;
    }
    return r;
}
// 32x32 Cholesky of the block in Ld (stride LDP) by one warp, right-looking, row `lane` in registers.
// The FP64 pipe is shared with the DMMAs of the co-resident CTA, so what matters is the length of
// the dependent FP64 chain: per column it is one shuffle, one rsqrt (no sqrt + divide), one multiply
// and one FMA; the 496 trailing updates are independent.  Writes L (upper part zeroed) back to Ld and
// the reciprocal diagonal to invd.  Returns the 1-based index of the first non-positive pivot, or 0.
__device__ int potrf_warp(double* Ld, double* invd, int lane) {
    double r[NB];
#pragma unroll
    for (int k = 0; k < NB; ++k) r[k] = Ld[lane * LDP + k];
    int fail = 0;
#pragma unroll
    for (int k = 0; k < NB; ++k) {
        const double piv = __shfl_sync(0xffffffffu, r[k], k);
        if (!(piv > 0.0) && fail == 0) fail = k + 1;
        const double inv = rsqrt(piv);
        const double l = lane > k ? r[k] * inv : (lane == k ? piv * inv : 0.0);
        r[k] = l;
        if (lane == k) invd[k] = inv;
#pragma unroll
        for (int j = k + 1; j < NB; ++j) {
            const double lj = __shfl_sync(0xffffffffu, l, j);
            r[j] = fma(-l, lj, r[j]);
        }
    }
#pragma unroll
    for (int k = 0; k < NB; ++k) Ld[lane * LDP + k] = k <= lane ? r[k] : 0.0;
    return fail;
}

template <int CH_WARPS>
__global__ void __launch_bounds__(CH_WARPS * 32, 16 / CH_WARPS) chol_solve_kernel(SolveParams p) {
    constexpr int CH_THREADS = CH_WARPS * 32;
    extern __shared__ double sm[];
    double* Ld = sm;                       // [32][33]
    double* invd = Ld + NB * LDP;          // [32] reciprocal diagonal of the current block (+ padding)
    double* red = invd + NB * LDP;         // [CH_WARPS][32]
    double* scratch = red + CH_WARPS * NB; // [40]
    double* xs = scratch + 40;             // [Nr]
    __shared__ int fail_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, tig = lane & 3;
    const int N = p.n_assets;
    const int ld = p.ldS;
    const int Nr = (N + NB - 1) / NB * NB;     // row index of the right-hand side

    for (int w = blockIdx.x; w < p.n_windows; w += gridDim.x) {
        double* S = p.S + (long long)w * p.win_stride;
        const double* rhs = p.rhs + (long long)w * p.ldv;
        if (tid == 0) fail_s = 0;
        for (int j = tid; j < ld; j += CH_THREADS) S[(long long)Nr * ld + j] = j < N ? rhs[j] : 0.0;
        __syncthreads();

        for (int j0 = 0; j0 < N; j0 += NB) {
            const int mt_total = (Nr + 8 - j0) / 8;        // m-tiles covering rows j0 .. Nr+7
            // ---------------- U: panel update (4 m-tiles of this warp at a time)
            for (int qb = warp; qb < mt_total; qb += 4 * CH_WARPS) {
                double acc[4][4][2];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) acc[i][nt][0] = acc[i][nt][1] = 0.0;
                int rowA[4];
                bool realA[4];
                int ni = 0;                  // m-tiles this warp really has in this pass (warp-uniform)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int q = qb + i * CH_WARPS;
                    rowA[i] = j0 + 8 * q + g;
                    realA[i] = q < mt_total && (rowA[i] < N || rowA[i] == Nr);
                    ni += q < mt_total ? 1 : 0;
                }
                for (int k0 = 0; k0 < j0; k0 += 16) {
                    d4 a[4], b[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        a[i] = load4(S + (long long)rowA[i] * ld + k0 + 2 * tig, realA[i]);
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) {
                        const int rb = j0 + 8 * nt + g;
                        b[nt] = load4(S + (long long)rb * ld + k0 + 2 * tig, rb < N);
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (i < ni) {        // skip the DMMAs of m-tile slots past the end of the panel
#pragma unroll
                            for (int qq = 0; qq < 4; ++qq)
#pragma unroll
                                for (int nt = 0; nt < 4; ++nt)
                                    dmma884(acc[i][nt][0], acc[i][nt][1], a[i].v[qq], b[nt].v[qq]);
                        }
                    }
                }
                // C = S - acc, masked for padding; diagonal-block tiles go to shared memory
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int q = qb + i * CH_WARPS;
                    if (q >= mt_total) continue;
                    const int row = rowA[i];
                    const bool in_diag = q < NB / 8;
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) {
                        const int col = j0 + 8 * nt + 2 * tig;
                        double c0, c1;
                        if (realA[i]) {
                            double2 s = make_double2(0.0, 0.0);
                            if (col < N) s = *reinterpret_cast<const double2*>(S + (long long)row * ld + col);
                            c0 = col < N ? s.x - acc[i][nt][0] : 0.0;
                            c1 = col + 1 < N ? s.y - acc[i][nt][1] : 0.0;
                        } else {
                            c0 = row == col ? 1.0 : 0.0;
                            c1 = row == col + 1 ? 1.0 : 0.0;
                        }
                        if (in_diag) {
                            const int lr = 8 * q + g, lc = 8 * nt + 2 * tig;
                            Ld[lr * LDP + lc] = c0;
                            Ld[lr * LDP + lc + 1] = c1;
                        } else if (realA[i] && col < N) {
                            *reinterpret_cast<double2*>(S + (long long)row * ld + col) = make_double2(c0, c1);
                        }
                    }
                }
            }
            __syncthreads();
            // ---------------- F: diagonal block
            if (warp == 0) {
                const int f = potrf_warp(Ld, invd, lane);
                if (lane == 0 && f != 0 && fail_s == 0) fail_s = j0 + f;
            }
            __syncthreads();
            // write L_d back (lower part, real rows only)
            for (int i = warp; i < NB; i += CH_WARPS) {
                const int row = j0 + i, col = j0 + lane;
                if (row < N && col < N && lane <= i) S[(long long)row * ld + col] = Ld[i * LDP + lane];
            }
            // ---------------- T: rows below the diagonal block (and the right-hand-side row):
            // x L_d' = c, one thread per row, right-looking so that the 496 updates are independent
            {
                const int nbelow = N - (j0 + NB) > 0 ? N - (j0 + NB) : 0;
                for (int e = tid; e < nbelow + 1; e += CH_THREADS) {
                    const int row = e < nbelow ? j0 + NB + e : Nr;
                    double* prow = S + (long long)row * ld + j0;
                    double pv[NB];
#pragma unroll
                    for (int c = 0; c < NB; c += 2) {
                        double2 v = make_double2(0.0, 0.0);
                        if (j0 + c < N) v = *reinterpret_cast<const double2*>(prow + c);
                        pv[c] = v.x;
                        pv[c + 1] = v.y;
                    }
#pragma unroll
                    for (int c = 0; c < NB; ++c) {
                        const double x = pv[c] * invd[c];
                        pv[c] = x;
#pragma unroll
                        for (int j = c + 1; j < NB; ++j) pv[j] = fma(-x, Ld[j * LDP + c], pv[j]);
                    }
#pragma unroll
                    for (int c = 0; c < NB; c += 2)
                        if (j0 + c < N) *reinterpret_cast<double2*>(prow + c) = make_double2(pv[c], pv[c + 1]);
                }
            }
            __syncthreads();
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
            if (col < N)
                for (int i = j0 + NB + warp; i < N; i += CH_WARPS)
                    part = fma(S[(long long)i * ld + col], xs[i], part);
            red[warp * NB + lane] = part;
            for (int i = warp; i < NB; i += CH_WARPS) {
                const int row = j0 + i;
                double v;
                if (row < N && col < N) v = lane <= i ? S[(long long)row * ld + col] : 0.0;
                else v = i == lane ? 1.0 : 0.0;
                Ld[i * LDP + lane] = v;
            }
            __syncthreads();
            if (warp == 0) {
                double r = xs[col];
#pragma unroll
                for (int wv = 0; wv < CH_WARPS; ++wv) r -= red[wv * NB + lane];
                const double rd = 1.0 / Ld[lane * LDP + lane];      // all 32 reciprocals in parallel
#pragma unroll
                for (int k = NB - 1; k >= 0; --k) {
                    const double xk = __shfl_sync(0xffffffffu, r * rd, k);
                    if (lane == k) r = xk;
                    if (lane < k) r = fma(-Ld[k * LDP + lane], xk, r);
                }
                xs[col] = r;
            }
            __syncthreads();
        }

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
    }
}

template <int CH_WARPS>
static cudaError_t launch_chol_t(const SolveParams& p, int sm_count, cudaStream_t st) {
    const int Nr = (p.n_assets + NB - 1) / NB * NB;
    const size_t smem = sizeof(double) * (size_t)(2 * NB * LDP + CH_WARPS * NB + 40 + Nr);
    cudaError_t e = cudaFuncSetAttribute(chol_solve_kernel<CH_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int grid = (16 / CH_WARPS) * sm_count;
    if (grid > p.n_windows) grid = p.n_windows;
    chol_solve_kernel<CH_WARPS><<<grid, CH_WARPS * 32, smem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_chol_solve(const SolveParams& p, int sm_count, cudaStream_t st) {
    if (p.n_windows <= 0) return cudaSuccess;
    static int warps = [] {
        const char* e = getenv("BP_CHOL_WARPS");      // tuning knob: warps per window (CTA), 16/warps CTAs per SM
        const int v = e ? atoi(e) : 4;
        return (v == 2 || v == 4 || v == 8) ? v : 4;
    }();
    if (warps == 2) return launch_chol_t<2>(p, sm_count, st);
    if (warps == 4) return launch_chol_t<4>(p, sm_count, st);
    return launch_chol_t<8>(p, sm_count, st);
}

}  // namespace bp
