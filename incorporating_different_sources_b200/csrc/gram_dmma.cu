// Batched windowed Gram / covariance kernel: FP64 tensor cores (DMMA.8x8x4) fed by TMA.
//
// For every rebalance window w and every lower-triangular 128x128 output tile (ti >= tj):
//
//   C  = alpha_w * R0[A rows, I]' R0[A rows, J]           phase A: intraday log returns (HF prior,
//                                                          portfolio_calculations.py:314-318)
//      +           R1[B rows, I]' R1[B rows, J]           phase B: daily log returns (:180-182)
//   out_ij = C_ij - p_i - p_j - beta_w * g_i * g_j         epilogue: rank-2 risk-free correction (:48-57),
//                                                          HF demeaning (:317) or Jeffreys t t'/n (:600-601)
//
// so one launch produces T, S0, S1 = S0 + T (:358) or the Jeffreys matrix directly; S0 and T are
// never materialised on the fused path.  R0 / R1 are the *shared* log-return matrices of the whole
// backtest: a window is a row range.
//
// Overlapping windows share almost all of their rows (consecutive rebalance dates differ by one daily
// row / one day of intraday bars), so the row range of a phase is split on a fixed block grid:
//   * rows inside whole blocks are NOT contracted again: the block's 128x128 Gram tile was computed once
//     (same kernel, tile_store_out mode) and is streamed from L2 into the accumulators ("ADD" items,
//     64 KB halves by cp.async.bulk into the same stage ring); two grid levels (coarse blocks in the
//     middle of the window, fine blocks next to its ends) keep both the number of tiles added and the
//     number of rows left for the tensor cores small;
//   * only the partial head / tail rows of the window go through the tensor cores ("K" items).
// All terms are still exact FP64 sums of the same products, only associated differently (no subtraction,
// no running update), so the result differs from a from-scratch contraction by a few ulp.
//
// Data movement: 3-D tensor maps (16-column group, row, group) with SWIZZLE_128B; one
// cp.async.bulk.tensor per 32-row x 128-column operand tile, 3-stage mbarrier ring that runs ahead
// across job boundaries (the next job's items are in flight during the epilogue).
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
constexpr int STAGE_BYTES = 2 * TILE_BYTES;          // A tile + B tile (= half of a stored block tile)
constexpr int GRAM_SMEM = GRAM_STAGES * STAGE_BYTES + 1024;
static_assert(STAGE_BYTES == GRAM_BLOCK_TILE_DOUBLES * 8 / 2, "an ADD item is half of a stored tile");

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// One job = (window, tile pair).  Its items, in order, per phase (A then B):
//   K items of the head rows, K items of the tail rows, ADD items (2 per block) of the coarse blocks, of the
//   fine blocks on the head side and of the fine blocks on the tail side
constexpr int NGROUPS = 10;
struct JobState {
    int job, w, ti, tj, pair;
    int base[NGROUPS];    // K groups: first row; ADD groups: first block
    int rows[NGROUPS];    // K groups: row count (ADD groups: unused)
    int end[NGROUPS];     // cumulative item counts
};
// group kinds per phase: 0,1 = K ; 2 = coarse ADD ; 3,4 = fine ADD
__device__ __forceinline__ bool group_is_add(int g) { return (g % 5) >= 2; }

__device__ __forceinline__ void job_setup(JobState& js, const GramParams& p, int job, int npairs) {
    js.job = job;
    const int wq = job / npairs;
    const int pr = job - wq * npairs;
    const int w = p.w_stride > 1 ? wq * p.w_stride : wq;
    int ti = 0;
    while ((ti + 1) * (ti + 2) / 2 <= pr) ++ti;
    js.w = w;
    js.ti = ti;
    js.tj = pr - ti * (ti + 1) / 2;
    js.pair = pr;
    const int* d = p.desc + (long long)w * GRAM_DESC_INTS;
    int e = 0;
#pragma unroll
    for (int ph = 0; ph < 2; ++ph) {
        const bool on = ph == 0 ? p.use_phaseA != 0 : p.use_phaseB != 0;
        const int* dp = d + ph * GRAM_PHASE_INTS;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const int g = ph * 5 + k;
            const int cnt = on ? dp[2 * k + 1] : 0;
            js.base[g] = dp[2 * k];
            js.rows[g] = cnt;
            e += k < 2 ? (cnt + GRAM_KT - 1) / GRAM_KT : 2 * cnt;
            js.end[g] = e;
        }
    }
}

// decode flat item f of a job: group and index inside the group
__device__ __forceinline__ void item_decode(const JobState& js, int f, int& grp, int& idx, int& base, int& rows) {
    grp = 0;
#pragma unroll
    for (int k = 0; k < NGROUPS - 1; ++k)
        if (f >= js.end[k]) grp = k + 1;
    int prev = 0;
    base = js.base[0];
    rows = js.rows[0];
#pragma unroll
    for (int k = 1; k < NGROUPS; ++k)
        if (grp == k) {
            prev = js.end[k - 1];
            base = js.base[k];
            rows = js.rows[k];
        }
    idx = f - prev;
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

// ADD item: half H of a stored tile in fragment-major order (register pair q of thread t at (q*256+t)*16 B):
// conflict-free LDS.128, every warp takes part
template <int H>
__device__ __forceinline__ void add_half(const unsigned char* s, double (&acc)[8][4][2], int tid) {
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const double2 v = *reinterpret_cast<const double2*>(s + (q * GRAM_THREADS + tid) * 16);
        acc[H * 4 + (q >> 2)][q & 3][0] += v.x;
        acc[H * 4 + (q >> 2)][q & 3][1] += v.y;
    }
}

// MV: also emit the G w0 by-product of phase A (a second instantiation, so that the plain kernel keeps its registers)
template <bool MV>
__global__ void __launch_bounds__(GRAM_THREADS, 1)
gram_dmma_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
                 const GramParams p) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t full_bar[GRAM_STAGES];
    __shared__ uint64_t empty_bar[GRAM_STAGES];    // one arrival per consumer warp: no block barrier per item
    __shared__ double mv_red[MV ? 6 : 1][GRAM_TILE];   // cross-warp partials of the optional G w0 by-product
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
        for (int s = 0; s < GRAM_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], GRAM_THREADS / 32);
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    __syncthreads();

    // per-thread fragment addressing (see header comment)
    const int offs0 = (2 * tig) * 128 + (((g >> 1) ^ (2 * tig)) << 4) + (g & 1) * 8;
    const int offs1 = (2 * tig + 1) * 128 + (((g >> 1) ^ (2 * tig + 1)) << 4) + (g & 1) * 8;
    const int warp_m0 = (warp & 1) * 64, warp_n0 = (warp >> 1) * 32;
    const int cgA0 = warp_m0 >> 4, cgB0 = warp_n0 >> 4;

    // ---- producer state (thread 0 only): runs GRAM_STAGES items ahead of the consumers
    JobState pj = {};
    int pf = 0;          // flat item index inside pj
    int pit = 0;         // items issued so far by this CTA
    bool pvalid = false;
    auto producer_advance_job = [&](int first) {
        int nj = first;                      // jobs without items are skipped by producer and consumer alike
        while (nj < njobs) {
            job_setup(pj, p, nj, npairs);
            if (pj.end[NGROUPS - 1] > 0) break;
            nj += gridDim.x;
        }
        pvalid = nj < njobs;
        pf = 0;
    };
    auto producer_issue = [&]() {
        const int stage = pit % GRAM_STAGES;
        unsigned char* sA = smem + stage * STAGE_BYTES;
        // the stage is free once every consumer warp has released its previous use
        if (pit >= GRAM_STAGES) mbar_wait(&empty_bar[stage], ((pit / GRAM_STAGES) - 1) & 1);
        int grp, idx, base, rows;
        item_decode(pj, pf, grp, idx, base, rows);
        if (group_is_add(grp)) {
            const double* store = p.store[grp / 5][(grp % 5) == 2 ? 0 : 1];
            const double* src = store + ((long long)(base + (idx >> 1)) * npairs + pj.pair) * GRAM_BLOCK_TILE_DOUBLES +
                                (idx & 1) * (GRAM_BLOCK_TILE_DOUBLES / 2);
            mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
            bulk_load(sA, src, STAGE_BYTES, &full_bar[stage]);
        } else {
            const int row = base + idx * GRAM_KT;
            const void* map = grp < 5 ? static_cast<const void*>(&map0) : static_cast<const void*>(&map1);
            const bool diag = pj.ti == pj.tj;
            mbar_arrive_expect_tx(&full_bar[stage], diag ? TILE_BYTES : STAGE_BYTES);
            tma_load_3d(sA, map, 0, row, pj.ti * 8, &full_bar[stage]);
            if (!diag) tma_load_3d(sA + TILE_BYTES, map, 0, row, pj.tj * 8, &full_bar[stage]);
        }
        ++pit;
        if (++pf == pj.end[NGROUPS - 1]) producer_advance_job(pj.job + gridDim.x);
    };
    if (tid == 0) {
        producer_advance_job(blockIdx.x);
        for (int s = 0; s < GRAM_STAGES && pvalid; ++s) producer_issue();
    }

    int it = 0;   // items consumed so far by this CTA
    JobState js;
    for (int job = blockIdx.x; job < njobs; job += gridDim.x) {
        job_setup(js, p, job, npairs);
        double acc[8][4][2];
#pragma unroll
        for (int mt = 0; mt < 8; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

        const bool diag = js.ti == js.tj;
        const int nitems = js.end[NGROUPS - 1];
        for (int f = 0; f < nitems; ++f, ++it) {
            const int stage = it % GRAM_STAGES;
            const uint32_t parity = (it / GRAM_STAGES) & 1;
            int grp, idx, base, rows;
            item_decode(js, f, grp, idx, base, rows);
            const unsigned char* sA = smem + stage * STAGE_BYTES;
            mbar_wait(&full_bar[stage], parity);
            if (group_is_add(grp)) {
                if (idx & 1) add_half<1>(sA, acc, tid);
                else add_half<0>(sA, acc, tid);
            } else {
                const int kvalid = min(GRAM_KT, rows - idx * GRAM_KT);
                const unsigned char* sB = diag ? sA : sA + TILE_BYTES;
                if (kvalid == GRAM_KT)
                    compute_tile<false>(sA, sB, kvalid, acc, offs0, offs1, cgA0, cgB0, tig);
                else
                    compute_tile<true>(sA, sB, kvalid, acc, offs0, offs1, cgA0, cgB0, tig);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[stage]);      // this warp is done with the stage
            if (tid == 0 && pvalid) producer_issue();
            if (MV && f == js.end[4] - 1) {
                // by-product of phase A: the raw intraday Gram tile times the prior weights (S0 w0 without a pass over
                // the window's own rows).  Row direction: y_i += sum_j G_ij w0_j; column direction (off-diagonal
                // tiles stand for their mirror image too): y_j += sum_i G_ij w0_i.  Partials per tile, fixed order.
                const double* w0 = p.mv_w0 + (long long)js.w * p.ldv;
                const int jb = js.tj * GRAM_TILE + warp_n0 + 2 * tig;
                const int ib = js.ti * GRAM_TILE + warp_m0 + g;
                double wj[4][2];
#pragma unroll
                for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int j = jb + nt * 8 + e;
                        wj[nt][e] = j < p.ldv ? w0[j] : 0.0;
                    }
                __syncthreads();                           // the previous job's readers of mv_red are done
#pragma unroll
                for (int mt = 0; mt < 8; ++mt) {
                    double r = 0.0;
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) r = fma(acc[mt][nt][0], wj[nt][0], fma(acc[mt][nt][1], wj[nt][1], r));
                    r += __shfl_xor_sync(0xffffffffu, r, 1);
                    r += __shfl_xor_sync(0xffffffffu, r, 2);
                    if (tig == 0) mv_red[warp >> 1][warp_m0 + mt * 8 + g] = r;
                }
                if (!diag) {
                    double wi[8];
#pragma unroll
                    for (int mt = 0; mt < 8; ++mt) {
                        const int i = ib + mt * 8;
                        wi[mt] = i < p.ldv ? w0[i] : 0.0;
                    }
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            double c = 0.0;
#pragma unroll
                            for (int mt = 0; mt < 8; ++mt) c = fma(acc[mt][nt][e], wi[mt], c);
                            c += __shfl_xor_sync(0xffffffffu, c, 4);
                            c += __shfl_xor_sync(0xffffffffu, c, 8);
                            c += __shfl_xor_sync(0xffffffffu, c, 16);
                            if (g == 0) mv_red[4 + (warp & 1)][warp_n0 + nt * 8 + 2 * tig + e] = c;
                        }
                }
                __syncthreads();
                double* dst = p.mv_part + ((long long)js.w * npairs + js.pair) * (2 * GRAM_TILE);
                if (tid < GRAM_TILE) dst[tid] = (mv_red[0][tid] + mv_red[1][tid]) + (mv_red[2][tid] + mv_red[3][tid]);
                else dst[tid] = diag ? 0.0 : mv_red[4][tid - GRAM_TILE] + mv_red[5][tid - GRAM_TILE];
            }
            if (f == js.end[4] - 1 && p.use_alpha) {      // last item of phase A: scale the HF Gram
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

        if (p.tile_store_out) {
            // block precompute: raw accumulators, fragment-major, fully coalesced 16-byte stores
            double* o = p.out + ((long long)js.w * npairs + js.pair) * GRAM_BLOCK_TILE_DOUBLES;
#pragma unroll
            for (int mt = 0; mt < 8; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt)
                    *reinterpret_cast<double2*>(o + ((mt * 4 + nt) * GRAM_THREADS + tid) * 2) =
                        make_double2(acc[mt][nt][0], acc[mt][nt][1]);
            continue;
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
        cudaError_t e = cudaFuncSetAttribute(gram_dmma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GRAM_SMEM);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(gram_dmma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GRAM_SMEM);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int nt = (p.n_assets + GRAM_TILE - 1) / GRAM_TILE;
    const long long njobs = (long long)p.n_windows * (nt * (nt + 1) / 2);
    const int grid = (int)(njobs < sm_count ? njobs : sm_count);
    if (p.mv_part && p.use_phaseA && !p.tile_store_out)
        gram_dmma_kernel<true><<<grid, GRAM_THREADS, GRAM_SMEM, st>>>(map0, map1, p);
    else
        gram_dmma_kernel<false><<<grid, GRAM_THREADS, GRAM_SMEM, st>>>(map0, map1, p);
    return cudaGetLastError();
}

}  // namespace bp
