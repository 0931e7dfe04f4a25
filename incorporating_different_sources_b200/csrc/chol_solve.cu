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
//   F  diagonal block 32x32 Cholesky + triangular inverse by one warp (row in registers,
//                     broadcasts through shared memory, warp shuffles for the pivot)
//   T  panel solve    L[j0+32:, j0:j0+32] = C * inv(L_d)'                       DMMA
// The right-hand side rides along as one extra row (row Nr = roundup(N,32)) of the matrix, so the
// forward substitution L z = b is a by-product of the factorisation.  The back substitution then
// walks the panels in reverse with coalesced row reads and a warp-level triangular solve.
//
// k-permutation: a lane (g, tig) loads the four consecutive doubles k0+4*tig .. k0+4*tig+3 of its
// row; DMMA step q (0..3) contracts the k-set {k0 + 4*tig + q}.  A and B fragments use the same
// assignment, so the contraction over the 16-wide chunk is exact while every global load is a
// full 32-byte sector.
#include "common.cuh"
#include "kernels.h"

namespace bp {

constexpr int CH_THREADS = 256;
constexpr int CH_WARPS = CH_THREADS / 32;
constexpr int NB = 32;
constexpr int LDP = 33;    // padded shared-memory row stride of the 32x32 blocks

struct d4 {
    double v[4];
};

__device__ __forceinline__ d4 load4(const double* ptr, bool pred) {
    d4 r;
    if (pred) {
        const double2 lo = *reinterpret_cast<const double2*>(ptr);
        const double2 hi = *reinterpret_cast<const double2*>(ptr + 2);
        r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = hi.x; r.v[3] = hi.y;
    } else {
        r.v[0] = r.v[1] = r.v[2] = r.v[3] = 0.0;
    }
    return r;
}
// as load4, but elements at index >= nvalid (columns beyond N, never written) read as zero
__device__ __forceinline__ d4 load4_bounded(const double* ptr, bool pred, int nvalid) {
    d4 r = load4(ptr, pred && nvalid > 0);
#pragma unroll
    for (int e = 0; e < 4; ++e)
        if (e >= nvalid) r.v[e] = 0.0;
    return r;
}

// 32x32 Cholesky of the block in Ld (stride LDP) by one warp; writes L (upper part zeroed) back to
// Ld and inv(L) to Li.  Returns the 1-based index of the first non-positive pivot, or 0.
__device__ int potrf_trtri_warp(double* Ld, double* Li, int lane) {
    double r[NB];
#pragma unroll
    for (int k = 0; k < NB; ++k) r[k] = Ld[lane * LDP + k];
    int fail = 0;
#pragma unroll
    for (int k = 0; k < NB; ++k) {
        double s = r[k];
#pragma unroll
        for (int q = 0; q < k; ++q) s = fma(-r[q], Ld[k * LDP + q], s);
        const double piv = __shfl_sync(0xffffffffu, s, k);
        if (!(piv > 0.0) && fail == 0) fail = k + 1;
        const double lkk = sqrt(piv);
        const double inv = 1.0 / lkk;
        r[k] = lane == k ? lkk : (lane > k ? s * inv : 0.0);
        Ld[lane * LDP + k] = r[k];
        __syncwarp();
    }
    // inverse of the lower-triangular block: lane j owns column j of X = inv(L)
    double x[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        double s = lane == i ? 1.0 : 0.0;
#pragma unroll
        for (int q = 0; q < i; ++q) s = fma(-Ld[i * LDP + q], x[q], s);
        x[i] = i < lane ? 0.0 : s / Ld[i * LDP + i];
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) Li[i * LDP + lane] = x[i];
    return fail;
}

__global__ void __launch_bounds__(CH_THREADS, 2) chol_solve_kernel(SolveParams p) {
    extern __shared__ double sm[];
    double* Ld = sm;                       // [32][33]
    double* Li = Ld + NB * LDP;            // [32][33]
    double* red = Li + NB * LDP;           // [CH_WARPS][32]
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
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int q = qb + i * CH_WARPS;
                    rowA[i] = j0 + 8 * q + g;
                    realA[i] = q < mt_total && (rowA[i] < N || rowA[i] == Nr);
                }
                for (int k0 = 0; k0 < j0; k0 += 16) {
                    d4 a[4], b[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        a[i] = load4(S + (long long)rowA[i] * ld + k0 + 4 * tig, realA[i]);
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) {
                        const int rb = j0 + 8 * nt + g;
                        b[nt] = load4(S + (long long)rb * ld + k0 + 4 * tig, rb < N);
                    }
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq)
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int nt = 0; nt < 4; ++nt)
                                dmma884(acc[i][nt][0], acc[i][nt][1], a[i].v[qq], b[nt].v[qq]);
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
                const int f = potrf_trtri_warp(Ld, Li, lane);
                if (lane == 0 && f != 0 && fail_s == 0) fail_s = j0 + f;
            }
            __syncthreads();
            // write L_d back (lower part, real rows only)
            for (int i = warp; i < NB; i += CH_WARPS) {
                const int row = j0 + i, col = j0 + lane;
                if (row < N && col < N && lane <= i) S[(long long)row * ld + col] = Ld[i * LDP + lane];
            }
            // ---------------- T: rows below the diagonal block:  X = C * inv(L_d)'
            for (int qb = warp; qb < mt_total; qb += CH_WARPS) {
                if (qb < NB / 8) continue;      // the diagonal block itself
                const int row = j0 + 8 * qb + g;
                const bool real = row < N || row == Nr;
                double acc[4][2];
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) acc[nt][0] = acc[nt][1] = 0.0;
#pragma unroll
                for (int kc = 0; kc < 2; ++kc) {
                    const int c0 = j0 + 16 * kc + 4 * tig;
                    const d4 a = load4_bounded(S + (long long)row * ld + c0, real, N - c0);
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) {
                        const int k = 16 * kc + 4 * tig + qq;
#pragma unroll
                        for (int nt = 0; nt < 4; ++nt)
                            dmma884(acc[nt][0], acc[nt][1], a.v[qq], Li[(8 * nt + g) * LDP + k]);
                    }
                }
                __syncwarp();   // all lanes have read the C tile before it is overwritten
                if (real) {
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) {
                        const int col = j0 + 8 * nt + 2 * tig;
                        if (col < N)
                            *reinterpret_cast<double2*>(S + (long long)row * ld + col) = make_double2(acc[nt][0], acc[nt][1]);
                    }
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
#pragma unroll
                for (int k = NB - 1; k >= 0; --k) {
                    const double xk = __shfl_sync(0xffffffffu, r, k) / Ld[k * LDP + k];
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

cudaError_t launch_chol_solve(const SolveParams& p, int sm_count, cudaStream_t st) {
    if (p.n_windows <= 0) return cudaSuccess;
    const int Nr = (p.n_assets + NB - 1) / NB * NB;
    const size_t smem = sizeof(double) * (size_t)(2 * NB * LDP + CH_WARPS * NB + 40 + Nr);
    cudaError_t e = cudaFuncSetAttribute(chol_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int grid = 2 * sm_count;
    if (grid > p.n_windows) grid = p.n_windows;
    chol_solve_kernel<<<grid, CH_THREADS, smem, st>>>(p);
    return cudaGetLastError();
}

}  // namespace bp
