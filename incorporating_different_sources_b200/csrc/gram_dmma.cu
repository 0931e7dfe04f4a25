// Batched windowed Gram / covariance kernel: FP64 tensor cores (DMMA.8x8x4) fed by TMA.
//
// For every rebalance window w and every lower-triangular 128x128 output tile (ti >= tj):
//
//   C  = alpha_w * R0[r0:r0+k0, I]' R0[r0:r0+k0, J]      segment 0: intraday log returns (HF prior,
//                                                         portfolio_calculations.py:314-318)
//      +           R1[r1:r1+k1, I]' R1[r1:r1+k1, J]      segment 1: daily log returns (:180-182)
//   out_ij = C_ij - p_i - p_j - beta_w * g_i * g_j        epilogue: rank-2 risk-free correction (:48-57),
//                                                         HF demeaning (:317) or Jeffreys t t'/n (:600-601)
//
// so one launch produces T, S0, S1 = S0 + T (:358) or the Jeffreys matrix directly; S0 and T are
// never materialised on the fused path.  R0 / R1 are the *shared* log-return matrices of the whole
// backtest: overlapping windows re-read the same rows through L2, HBM sees each row about once.
//
// Data movement: 3-D tensor maps (16-column group, row, group) with SWIZZLE_128B; one
// cp.async.bulk.tensor per 32-row x 128-column operand tile, 3-stage mbarrier ring that runs ahead
// across job boundaries (the next job's tiles are in flight during the epilogue).
// Math: 8 warps, 64x32 warp tiles, mma.sync.m8n8k4.f64.  The summation index inside an 8-row group
// is permuted (lane tig reads row 2*tig+s) which makes every LDS.64 fragment load bank-conflict
// free under the 128B swizzle; A and B use the same permutation so the product is unchanged.
#include "common.cuh"
#include "kernels.h"

namespace bp {

constexpr int GRAM_THREADS = 256;
constexpr int GRAM_STAGES = 3;
constexpr int CG_BYTES = GRAM_KT * 128;              // one 16-column group: [KT rows][16 doubles]
constexpr int TILE_BYTES = 8 * CG_BYTES;             // 128 columns
constexpr int STAGE_BYTES = 2 * TILE_BYTES;          // A tile + B tile
constexpr int GRAM_SMEM = GRAM_STAGES * STAGE_BYTES + 1024;

struct JobState {
    int job;
    int w, ti, tj;
    int row0_0, row0_1;   // first row of segment 0 / 1
    int rows_0, rows_1;   // row count of segment 0 / 1
    int nkt0;     // k-tiles of segment 0
    int nkt;      // k-tiles of both segments
};

__device__ __forceinline__ void job_setup(JobState& js, const GramParams& p, int job, int npairs) {
    js.job = job;
    const int w = job / npairs;
    const int pr = job - w * npairs;
    int ti = 0;
    while ((ti + 1) * (ti + 2) / 2 <= pr) ++ti;
    js.w = w;
    js.ti = ti;
    js.tj = pr - ti * (ti + 1) / 2;
    if (p.seg0_row0) {
        js.row0_0 = p.seg0_row0[w] + p.seg0_row_bias;
        js.rows_0 = p.seg0_rows[w] + p.seg0_rows_bias;
    } else {
        js.row0_0 = 0;
        js.rows_0 = 0;
    }
    if (p.seg1_row0) {
        js.row0_1 = p.seg1_row0[w] + p.seg1_row_bias;
        js.rows_1 = p.seg1_rows ? p.seg1_rows[w] : p.seg1_rows_const;
    } else {
        js.row0_1 = 0;
        js.rows_1 = 0;
    }
    js.nkt0 = (js.rows_0 + GRAM_KT - 1) / GRAM_KT;
    js.nkt = js.nkt0 + (js.rows_1 + GRAM_KT - 1) / GRAM_KT;
}

template <bool MASK>
__device__ __forceinline__ void compute_tile(const unsigned char* sA, const unsigned char* sB, int kvalid,
                                             double (&acc)[8][4][2], int offs0, int offs1, int cgA0, int cgB0,
                                             int tig) {
    const int ngroups = MASK ? (kvalid + 7) >> 3 : GRAM_KT / 8;
#pragma unroll 1
    for (int r8 = 0; r8 < ngroups; ++r8) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int offs = (s ? offs1 : offs0) + r8 * 1024;
            double a[8], b[4];
#pragma unroll
            for (int mt = 0; mt < 8; ++mt)
                a[mt] = *reinterpret_cast<const double*>(sA + (cgA0 + (mt >> 1)) * CG_BYTES + (offs ^ ((mt & 1) << 6)));
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
                b[nt] = *reinterpret_cast<const double*>(sB + (cgB0 + (nt >> 1)) * CG_BYTES + (offs ^ ((nt & 1) << 6)));
            if (MASK) {
                if (r8 * 8 + 2 * tig + s >= kvalid) {
#pragma unroll
                    for (int mt = 0; mt < 8; ++mt) a[mt] = 0.0;
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) b[nt] = 0.0;
                }
            }
#pragma unroll
            for (int mt = 0; mt < 8; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
        }
    }
}

__global__ void __launch_bounds__(GRAM_THREADS, 1)
gram_dmma_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
                 const GramParams p) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t full_bar[GRAM_STAGES];
    // 1024-byte alignment for the 128B swizzle, by pointer arithmetic on the shared array itself so that
    // the compiler keeps the shared address space (an integer round trip degrades every LDS to a generic LD)
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, tig = lane & 3;
    const int nt_tiles = (p.n_assets + GRAM_TILE - 1) / GRAM_TILE;
    const int npairs = nt_tiles * (nt_tiles + 1) / 2;
    const int njobs = p.n_windows * npairs;

    if (tid == 0) {
        tma_prefetch_desc(&map0);
        tma_prefetch_desc(&map1);
        for (int s = 0; s < GRAM_STAGES; ++s) mbar_init(&full_bar[s], 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    __syncthreads();

    // per-thread fragment addressing (see header comment)
    const int offs0 = (2 * tig) * 128 + (((g >> 1) ^ (2 * tig)) << 4) + (g & 1) * 8;
    const int offs1 = (2 * tig + 1) * 128 + (((g >> 1) ^ (2 * tig + 1)) << 4) + (g & 1) * 8;
    const int warp_m0 = (warp & 1) * 64, warp_n0 = (warp >> 1) * 32;
    const int cgA0 = warp_m0 >> 4, cgB0 = warp_n0 >> 4;

    // ---- producer state (thread 0 only): runs GRAM_STAGES tiles ahead of the consumers
    JobState pj = {};
    int pf = 0;          // flat k-tile index inside pj
    int pit = 0;         // tiles issued so far by this CTA
    bool pvalid = false;
    auto producer_issue = [&]() {
        // issue tile (pj, pf) into stage pit % STAGES, then advance
        const int stage = pit % GRAM_STAGES;
        unsigned char* sA = smem + stage * STAGE_BYTES;
        const int seg = pf < pj.nkt0 ? 0 : 1;
        const int kt = seg ? pf - pj.nkt0 : pf;
        const int row = (seg ? pj.row0_1 : pj.row0_0) + kt * GRAM_KT;
        const void* map = seg ? static_cast<const void*>(&map1) : static_cast<const void*>(&map0);
        const bool diag = pj.ti == pj.tj;
        mbar_arrive_expect_tx(&full_bar[stage], diag ? TILE_BYTES : STAGE_BYTES);
        tma_load_3d(sA, map, 0, row, pj.ti * 8, &full_bar[stage]);
        if (!diag) tma_load_3d(sA + TILE_BYTES, map, 0, row, pj.tj * 8, &full_bar[stage]);
        ++pit;
        if (++pf == pj.nkt) {
            pf = 0;
            const int nj = pj.job + gridDim.x;
            if (nj < njobs) job_setup(pj, p, nj, npairs);
            else pvalid = false;
        }
    };
    if (tid == 0 && (int)blockIdx.x < njobs) {
        job_setup(pj, p, blockIdx.x, npairs);
        pvalid = true;
        for (int s = 0; s < GRAM_STAGES && pvalid; ++s) producer_issue();
    }

    int it = 0;   // tiles consumed so far by this CTA
    JobState js;
    for (int job = blockIdx.x; job < njobs; job += gridDim.x) {
        job_setup(js, p, job, npairs);
        double acc[8][4][2];
#pragma unroll
        for (int mt = 0; mt < 8; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

        const bool diag = js.ti == js.tj;
        for (int f = 0; f < js.nkt; ++f, ++it) {
            const int stage = it % GRAM_STAGES;
            const uint32_t parity = (it / GRAM_STAGES) & 1;
            const int seg = f < js.nkt0 ? 0 : 1;
            const int kt = seg ? f - js.nkt0 : f;
            const int kvalid = min(GRAM_KT, (seg ? js.rows_1 : js.rows_0) - kt * GRAM_KT);
            const unsigned char* sA = smem + stage * STAGE_BYTES;
            const unsigned char* sB = diag ? sA : sA + TILE_BYTES;
            mbar_wait(&full_bar[stage], parity);
            if (kvalid == GRAM_KT)
                compute_tile<false>(sA, sB, kvalid, acc, offs0, offs1, cgA0, cgB0, tig);
            else
                compute_tile<true>(sA, sB, kvalid, acc, offs0, offs1, cgA0, cgB0, tig);
            __syncthreads();                       // every warp is done with this stage
            if (tid == 0 && pvalid) producer_issue();
            if (seg == 0 && f == js.nkt0 - 1 && p.use_alpha) {
                const double alpha = p.scal[(long long)js.w * BP_S_COUNT + BP_S_ALPHA];
#pragma unroll
                for (int mt = 0; mt < 8; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) {
                        acc[mt][nt][0] *= alpha;
                        acc[mt][nt][1] *= alpha;
                    }
            }
        }

        // ---- epilogue: out_ij = C_ij - p_i - p_j - beta g_i g_j
        const int N = p.n_assets;
        const double beta = p.use_beta ? p.scal[(long long)js.w * BP_S_COUNT + BP_S_BETA] : 0.0;
        const double* pv = p.pvec ? p.pvec + (long long)js.w * p.ldv : nullptr;
        const double* gv = p.gvec ? p.gvec + (long long)js.w * p.ldv : nullptr;
        double* out = p.out + (long long)js.w * p.win_stride;
        const int i_base = js.ti * GRAM_TILE + warp_m0 + g;
        const int j_base = js.tj * GRAM_TILE + warp_n0 + 2 * tig;
        double pj_[4][2], gj_[4][2];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = j_base + nt * 8 + e;
                pj_[nt][e] = (pv && j < N) ? pv[j] : 0.0;
                gj_[nt][e] = (gv && j < N) ? gv[j] : 0.0;
            }
#pragma unroll
        for (int mt = 0; mt < 8; ++mt) {
            const int i = i_base + mt * 8;
            if (i >= N) continue;
            const double pi = pv ? pv[i] : 0.0;
            const double bgi = gv ? beta * gv[i] : 0.0;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int j = j_base + nt * 8;
                if (j >= N) continue;
                const double v0 = acc[mt][nt][0] - pi - pj_[nt][0] - bgi * gj_[nt][0];
                const double v1 = acc[mt][nt][1] - pi - pj_[nt][1] - bgi * gj_[nt][1];
                double* dst = out + (long long)i * p.ldS + j;
                if (j + 1 < N) {
                    *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
                } else {
                    dst[0] = v0;
                }
                if (p.mirror && !diag) {
                    out[(long long)j * p.ldS + i] = v0;
                    if (j + 1 < N) out[(long long)(j + 1) * p.ldS + i] = v1;
                }
            }
        }
    }
}

cudaError_t launch_gram(const GramParams& p, const CUtensorMap& map0, const CUtensorMap& map1, int sm_count,
                        cudaStream_t st) {
    if (p.n_windows <= 0) return cudaSuccess;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gram_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GRAM_SMEM);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int nt = (p.n_assets + GRAM_TILE - 1) / GRAM_TILE;
    const long long njobs = (long long)p.n_windows * (nt * (nt + 1) / 2);
    const int grid = (int)(njobs < sm_count ? njobs : sm_count);
    gram_dmma_kernel<<<grid, GRAM_THREADS, GRAM_SMEM, st>>>(map0, map1, p);
    return cudaGetLastError();
}

}  // namespace bp
