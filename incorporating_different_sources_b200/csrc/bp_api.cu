// C-ABI of libbayes_portfolio.so: market residency, window batches, stage orchestration.
// Declarations and the reference interfaces each entry point replaces: include/bayes_portfolio.h.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <array>
#include <cstdarg>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "../../include/bayes_portfolio.h"
#include "kernels.h"

using namespace bp;

static_assert(BP_NSCAL == bp::BP_S_COUNT, "scalar record size mismatch between header and kernels");

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU_TRY(expr)                                                                                     \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return fail(BP_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

bool is_device_ptr(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

}  // namespace

struct bp_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    int sm_count = 0;
    size_t ws_limit = (size_t)24 << 30;
    long long launches = 0;
    // market
    bool has_market = false;
    int N = 0, D = 0, ld = 0, n_mcm = 0;
    long long R = 0;
    // prices / caps / hf_prices keep the caller's dense [rows][N] layout; lr_d / lr_hf are padded to ld
    double *prices = nullptr, *lr_d = nullptr, *caps = nullptr, *hf_prices = nullptr, *lr_hf = nullptr,
           *mcm = nullptr, *rf_row = nullptr;
    bool has_caps = false;
    // resident POOL: the full market (every column, every row) kept in HBM so that the working market of a batch
    // (a column subset in a given order, a row range) is gathered on the device instead of uploaded again
    double *pool_prices = nullptr, *pool_caps = nullptr, *pool_hf = nullptr, *pool_rf = nullptr;
    int pool_N = 0, pool_D = 0;
    long long pool_R = 0;
    int* pool_cols = nullptr;
    size_t pool_cols_cap = 0;
    // resampled (weekly) return rows: see bp_set_resampled
    double *lr_w = nullptr, *rf_w = nullptr, *mcm_w = nullptr;
    int* rs_idx = nullptr;
    int Rw = 0;
    CUtensorMap map_w;
    size_t cap_daily = 0, cap_hf = 0, cap_mcm = 0, cap_days = 0;   // allocated sizes (elements) for buffer reuse
    cudaStream_t copy_stream = nullptr;                  // intraday H2D + its log-return kernel
    cudaEvent_t ev_main = nullptr, ev_hf = nullptr;
    bool hf_pending = false;
    // an asynchronous intraday upload is cut into segments (copy + log returns + event each) so that the
    // conjugate stages of the early rebalance dates run while the later bars are still on the bus
    static constexpr int MAX_SEG = 8;
    int pipe_segments = MAX_SEG;
    size_t pipe_min_bytes = (size_t)256 << 20;
    int n_seg = 0, seg_waited = 0;
    int n_frac = 0;                            // > 0: explicit cumulative row fractions of the segments (bp_set_upload_fractions)
    double frac[MAX_SEG] = {0};
    long long lr_hf_done = 0;                  // intraday return rows < lr_hf_done have been computed
    long long hf_extra_rows = 0;               // rows allocated behind lr_hf for gathered overnight returns
    int lr_hf_ld_cap = 1;
    long long lr_hf_rows_cap = 0;              // rows allocated for lr_hf (at the leading dimension it was allocated with)
    long long seg_end[MAX_SEG] = {0};          // return rows < seg_end[s] are valid once ev_seg[s] has fired
    cudaEvent_t ev_seg[MAX_SEG] = {nullptr};
    CUtensorMap map_d, map_hf;
    // window descriptors on the device: day_row, span, row0, hf_row0, hf_m
    int* desc = nullptr;
    int* desc_host = nullptr;          // page-locked staging of the descriptors (read zero-copy by a kernel)
    size_t desc_cap = 0;               // capacity in ints
    // two descriptor slots used alternately: planning the next batch on the host never waits for the GPU to
    // have fetched (or finished using) the previous one, only for the batch before that
    int* desc_slot[2] = {nullptr, nullptr};
    int* desc_host_slot[2] = {nullptr, nullptr};
    size_t desc_cap_slot[2] = {0, 0};
    cudaEvent_t ev_desc[2] = {nullptr, nullptr};
    bool ev_desc_armed[2] = {false, false};
    int desc_turn = 0;
    bool async_outputs = false;        // host outputs are complete only after bp_synchronize (see bp_set_async_outputs)
    // block tile stores of the Gram kernel (window-overlap reuse) and the smallest batch that uses them
    double* store[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};      // [phase][level]
    size_t store_cap[2][2] = {{0, 0}, {0, 0}};                           // capacity in doubles
    double* rstore[2] = {nullptr, nullptr};                              // phase-B range-sum stores [level]
    size_t rstore_cap[2] = {0, 0};
    double* fstore = nullptr;                                            // phase-B combined runs (coarse + fine + fine)
    size_t fstore_cap = 0;
    int reuse_min_windows = 32;
    // pre-summed intraday day blocks (long HF look-backs): windows with at least this many trading days add <= 3 scanned
    // tiles instead of one tile per day (0 disables); vector store [3 nb][ld] of the scanned per-day column sums
    int presum_min_days = 8;
    double* vstore = nullptr;
    size_t vstore_cap = 0;
    double work_scan_tiles = 0;
    // work counters of the Gram stage since the last bp_get_gram_work (bench.py's roofline accounting)
    double work_k_rows = 0, work_add_blocks = 0, work_pre_rows = 0, work_full_rows = 0;
    // Jeffreys windows of consecutive trade dates: only every chain_group-th window is factorised, the others are
    // solved relative to it (jeffreys_chain.cu); < 2 disables.  Work counters for bench.py's roofline accounting.
    int chain_group = 8;
    int work_stride = 1;               // upload_batch counts the Gram work of every work_stride-th window only
    double work_factored = 0, work_chained = 0;
    double* prior_n = nullptr;
    int prior_n_cap = 0;
    // banded-GEMM daily pass (band_prep.cu): risk-free weights [W][band_ld] and their sums [W][2]
    double *band_aw = nullptr, *band_stats = nullptr;
    size_t band_aw_cap = 0, band_stats_cap = 0;
    // workspace
    unsigned char* ws = nullptr;
    size_t ws_bytes = 0;
    unsigned char* stage = nullptr;
    size_t stage_bytes = 0;
    bool need_sync = false;
    PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
    // optional per-stage CUDA-event timing (bench.py's roofline numbers)
    bool timing = false;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    struct Span { int stage; cudaEvent_t a, b; };
    std::vector<Span> spans;
    cudaEvent_t ev_t0 = nullptr;       // start of the last upload (BP_TIMELINE=1: spans are printed relative to it)
    bool t0_armed = false;
};

namespace {
struct StageTimer {
    bp_handle* h;
    int stage;
    cudaEvent_t a = nullptr, b = nullptr;
    StageTimer(bp_handle* h_, int stage_) : h(h_), stage(stage_) {
        if (!h->timing) return;
        while (h->ev_pool.size() < h->ev_used + 2) {
            cudaEvent_t e;
            if (cudaEventCreate(&e) != cudaSuccess) return;
            h->ev_pool.push_back(e);
        }
        a = h->ev_pool[h->ev_used++];
        b = h->ev_pool[h->ev_used++];
        cudaEventRecord(a, h->stream);
    }
    ~StageTimer() {
        if (!a) return;
        cudaEventRecord(b, h->stream);
        h->spans.push_back({stage, a, b});
    }
};
}  // namespace

namespace {

void free_resampled(bp_handle* h) {
    cudaFree(h->lr_w); cudaFree(h->rf_w); cudaFree(h->mcm_w); cudaFree(h->rs_idx);
    h->lr_w = h->rf_w = h->mcm_w = nullptr;
    h->rs_idx = nullptr;
    h->Rw = 0;
}

void free_pool(bp_handle* h) {
    cudaFree(h->pool_prices); cudaFree(h->pool_caps); cudaFree(h->pool_hf); cudaFree(h->pool_rf);
    h->pool_prices = h->pool_caps = h->pool_hf = h->pool_rf = nullptr;
    h->pool_N = h->pool_D = 0;
    h->pool_R = 0;
}

void free_market(bp_handle* h) {
    free_resampled(h);
    cudaFree(h->prices); cudaFree(h->lr_d); cudaFree(h->caps); cudaFree(h->hf_prices);
    cudaFree(h->lr_hf); cudaFree(h->mcm); cudaFree(h->rf_row);
    h->prices = h->lr_d = h->caps = h->hf_prices = h->lr_hf = h->mcm = h->rf_row = nullptr;
    h->cap_daily = h->cap_hf = h->cap_mcm = h->cap_days = 0;
    h->has_market = false;
}

// Conjugate stages read lr_hf.  The intraday prices arrive on the copy stream; their log returns are computed
// on the COMPUTE stream once the copy event has fired (a kernel on the copy stream would have to wait for
// free SMs behind the persistent Gram / solve kernels, and the next segment's copy would queue behind it).
int wait_hf(bp_handle* h) {
    if (h->hf_pending) {
        CU_TRY(cudaStreamWaitEvent(h->stream, h->ev_hf, 0));
        launch_log_returns(h->hf_prices, h->N, h->lr_hf, h->ld, h->R, h->N, h->sm_count, h->stream, h->lr_hf_done);
        h->launches++;
        CU_TRY(cudaGetLastError());
        h->lr_hf_done = h->R;
        h->hf_pending = false;
        h->seg_waited = h->n_seg;
    }
    return BP_OK;
}

// The same for a batch that reads intraday PRICE rows < rows_needed only: waits for the upload segments that cover them
// and computes their log returns; the later segments stay pending.  A caller that cuts a date-sorted batch into
// sub-batches along the segment boundaries thereby overlaps the compute of sub-batch k with the copy of segment k+1 on
// ANY path (the pipelined branch of run_batches does this inside one call, but only for the one-tile-per-day form).
int wait_hf_rows(bp_handle* h, long long rows_needed) {
    if (!h->hf_pending) return BP_OK;
    if (h->n_seg <= 1 || rows_needed >= h->R) return wait_hf(h);
    while (h->seg_waited < h->n_seg && h->lr_hf_done < rows_needed) {
        const int s = h->seg_waited;
        CU_TRY(cudaStreamWaitEvent(h->stream, h->ev_seg[s], 0));
        launch_log_returns(h->hf_prices, h->N, h->lr_hf, h->ld, h->seg_end[s], h->N, h->sm_count, h->stream, h->lr_hf_done);
        h->launches++;
        CU_TRY(cudaGetLastError());
        h->lr_hf_done = h->seg_end[s];
        h->seg_waited = s + 1;
    }
    if (h->seg_waited >= h->n_seg) {
        h->hf_pending = false;
        h->lr_hf_done = h->R;
    }
    return BP_OK;
}

int make_map(bp_handle* h, CUtensorMap* map, const double* base, long long rows, int ld) {
    // 3-D view (16-column group element, row, column group) of a row-major [rows][ld] matrix
    cuuint64_t dims[3] = {16, (cuuint64_t)rows, (cuuint64_t)(ld / 16)};
    cuuint64_t strides[2] = {(cuuint64_t)ld * sizeof(double), 16 * sizeof(double)};
    cuuint32_t box[3] = {16, (cuuint32_t)GRAM_KT, 8};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = h->encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), dims, strides, box,
                           estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(BP_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return BP_OK;
}

// 2-D tensor map over a solver workspace of `rows` x ldS doubles (32-row x 16-column boxes)
int make_solve_map(bp_handle* h, CUtensorMap* map, const double* base, long long rows, int ldS) {
    cuuint64_t dims[2] = {(cuuint64_t)ldS, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ldS * sizeof(double)};
    cuuint32_t box[2] = {16, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = h->encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box,
                           estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(BP_ERR_CUDA, "cuTensorMapEncodeTiled (solver workspace) failed with CUresult %d", (int)r);
    return BP_OK;
}

int ensure_ws(bp_handle* h, size_t bytes) {
    if (bytes <= h->ws_bytes) return BP_OK;
    if (h->ws) cudaFree(h->ws);
    h->ws = nullptr;
    h->ws_bytes = 0;
    CU_TRY(cudaMalloc(&h->ws, bytes));
    h->ws_bytes = bytes;
    return BP_OK;
}

int ensure_stage(bp_handle* h, size_t bytes) {
    if (bytes <= h->stage_bytes) return BP_OK;
    if (h->stage) {
        CU_TRY(cudaStreamSynchronize(h->stream));
        cudaFree(h->stage);
    }
    h->stage = nullptr;
    h->stage_bytes = 0;
    CU_TRY(cudaMalloc(&h->stage, bytes));
    h->stage_bytes = bytes;
    return BP_OK;
}

struct Chunk {
    // carved from the workspace for Wc windows
    double *S, *t, *pvec, *gvec, *rhs, *w0, *s0w0, *w1, *nu, *weights, *scal, *y, *mv;
    int* status;
};

struct Layout {
    int N, ldv, ldS, rowsS;
    long long win_stride, y_stride, mv_stride;
    size_t per_window;
};

// max_m: longest intraday window (scratch of the row dots); with_mv: room for the Gram kernel's G w0 partials
Layout make_layout(const bp_handle* h, int max_m, bool with_mv = false) {
    Layout L;
    L.N = h->N;
    L.ldv = h->ld;
    L.ldS = round_up(h->N, 32);
    L.rowsS = L.ldS + 8;
    L.win_stride = (long long)L.rowsS * L.ldS;
    L.y_stride = round_up(std::max(max_m, 2), 2);
    const int nt = (h->N + GRAM_TILE - 1) / GRAM_TILE;
    L.mv_stride = with_mv ? (long long)(nt * (nt + 1) / 2) * 2 * GRAM_TILE : 0;
    L.per_window = sizeof(double) * ((size_t)L.win_stride + 9 * (size_t)L.ldv + BP_NSCAL + (size_t)L.y_stride + (size_t)L.mv_stride) + 16;
    return L;
}

Chunk carve(unsigned char* ws, const Layout& L, int Wc) {
    Chunk c;
    double* p = reinterpret_cast<double*>(ws);
    c.S = p;        p += (size_t)Wc * L.win_stride;
    c.t = p;        p += (size_t)Wc * L.ldv;
    c.pvec = p;     p += (size_t)Wc * L.ldv;
    c.gvec = p;     p += (size_t)Wc * L.ldv;
    c.rhs = p;      p += (size_t)Wc * L.ldv;
    c.w0 = p;       p += (size_t)Wc * L.ldv;
    c.s0w0 = p;     p += (size_t)Wc * L.ldv;
    c.w1 = p;       p += (size_t)Wc * L.ldv;
    c.nu = p;       p += (size_t)Wc * L.ldv;
    c.weights = p;  p += (size_t)Wc * L.ldv;
    c.scal = p;     p += (size_t)Wc * BP_NSCAL;
    c.y = p;        p += (size_t)Wc * L.y_stride;
    c.mv = p;       p += (size_t)Wc * L.mv_stride;
    c.status = reinterpret_cast<int*>(p);
    return c;
}

// the carved arrays as seen from window k of the chunk
Chunk chunk_at(const Chunk& c, const Layout& L, int k) {
    Chunk o = c;
    o.S += (size_t)k * L.win_stride;
    o.t += (size_t)k * L.ldv;       o.pvec += (size_t)k * L.ldv;  o.gvec += (size_t)k * L.ldv;
    o.rhs += (size_t)k * L.ldv;     o.w0 += (size_t)k * L.ldv;    o.s0w0 += (size_t)k * L.ldv;
    o.w1 += (size_t)k * L.ldv;      o.nu += (size_t)k * L.ldv;    o.weights += (size_t)k * L.ldv;
    o.scal += (size_t)k * BP_NSCAL;
    o.y += (size_t)k * L.y_stride;
    o.mv += (size_t)k * L.mv_stride;
    o.status += k;
    return o;
}

// validated, device-resident description of one batch
struct Batch {
    int W = 0, n = 0, max_m = 0;
    bool has_hf = false;
    const int *day_row = nullptr, *span = nullptr, *row0 = nullptr, *hf_row0 = nullptr, *hf_m = nullptr;
    const double* prior_n = nullptr;
    int mcm_rows = 0;
    bool resampled = false;
    const int *extra_row = nullptr, *caps_row = nullptr;
    const int* gdesc = nullptr;        // [W][GRAM_DESC_INTS] Gram job descriptors
    const int* bdesc[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // block precompute launches [phase][level]
    int nblocks[2][2] = {{0, 0}, {0, 0}};
    const int* rdesc[2] = {nullptr, nullptr};      // phase B: [lo, hi) block ranges to pre-sum, per level
    int nranges[2] = {0, 0};
    const int* fdesc = nullptr;                    // phase B: [coarse, fineA, fineB] run ids to combine
    int nfull = 0;
    bool band_ok = false;              // consecutive trade dates: the daily pass runs as a banded GEMM
    // intraday block grid (for the pipelined upload): block k of level l ends at return row hf_off + (hf_bmin[l]+k+1)*hf_blk[l]
    int hf_off = 0, hf_bmin[2] = {0, 0}, hf_blk[2] = {0, 0};
    bool hf_inner = false;             // day blocks without their overnight return + gathered overnight rows (PhasePlan::inner)
    int hf_nb0 = 0;
    std::vector<int> hf_starts;        // inner: price row where day block k starts (nb + 1 entries), host copy
    const int* hf_starts_dev = nullptr;
    bool presum = false;               // windows add <= 3 scanned tiles (PhasePlan::chunk)
    int presum_chunk = 0;
    const int* hf_vids = nullptr;      // [W][3] scanned-vector ids of each window
};

// Block grids of one phase, two levels (0 = coarse, 1 = fine; the fine size divides the coarse size): block b of
// level l covers return rows [off + b*blk[l], off + (b+1)*blk[l]); blk[l] == 0: level unused
struct PhasePlan {
    int blk[2] = {0, 0};
    int off = 0;
    int bmin[2] = {0, 0}, nb[2] = {0, 0};
    // "inner" day blocks: every hf_lo / hf_hi of the batch is a block boundary (the HF look-back always covers whole
    // days, :310-312, whatever the bar calendar), block b = return rows starts[b]+1 .. starts[b+1]-1, i.e. a trading
    // day WITHOUT its first (overnight) return; the overnight rows are gathered into a compact row range behind the
    // intraday matrix, so a window is [its days' inner blocks] + [one short run of overnight rows]
    bool inner = false;
    std::vector<int> starts;
    // pre-summed runs (long look-backs): chunk > 0 -> a window adds suffix'[first day] + prefix[...] tiles (<= 3)
    int chunk = 0;
};

inline long long floor_div(long long a, long long b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }

// Split of the row range [row0, row0+rows) of a window: head rows, fine blocks, coarse blocks, fine blocks, tail rows
struct Split {
    long long k0_row = 0, k0_n = 0, k1_row = 0, k1_n = 0;     // rows contracted on the tensor cores
    long long c_lo = 0, c_hi = 0;                             // coarse blocks [c_lo, c_hi)
    long long fa_lo = 0, fa_hi = 0, fb_lo = 0, fb_hi = 0;     // fine blocks on the head / tail side
};

Split split_rows(int row0, int rows, const PhasePlan& P) {
    Split s;
    s.k0_row = row0;
    s.k0_n = rows;
    if (rows <= 0 || (P.blk[0] <= 0 && P.blk[1] <= 0)) return s;
    const long long r0 = (long long)row0 - P.off, r1 = r0 + rows;
    long long head_end = r1, tail_begin = r1;      // relative rows; head = [r0, head_end), tail = [tail_begin, r1)
    if (P.blk[0] > 0) {
        const long long B = P.blk[0];
        const long long lo = floor_div(r0 + B - 1, B), hi = floor_div(r1, B);
        if (hi > lo) {
            s.c_lo = lo;
            s.c_hi = hi;
            head_end = lo * B;
            tail_begin = hi * B;
        }
    }
    if (P.blk[1] > 0) {
        const long long F = P.blk[1];
        if (s.c_hi > s.c_lo) {
            const long long f_lo = floor_div(r0 + F - 1, F), f_hi = floor_div(r1, F);
            s.fa_lo = f_lo;  s.fa_hi = head_end / F;          // head_end is a multiple of the coarse (hence fine) size
            s.fb_lo = tail_begin / F;  s.fb_hi = f_hi;
            s.k0_n = f_lo * F - r0;
            s.k1_row = P.off + f_hi * F;
            s.k1_n = r1 - f_hi * F;
        } else {
            const long long f_lo = floor_div(r0 + F - 1, F), f_hi = floor_div(r1, F);
            if (f_hi > f_lo) {
                s.fa_lo = f_lo;  s.fa_hi = f_hi;
                s.k0_n = f_lo * F - r0;
                s.k1_row = P.off + f_hi * F;
                s.k1_n = r1 - f_hi * F;
            }
        }
    } else if (s.c_hi > s.c_lo) {
        s.k0_n = head_end - r0;
        s.k1_row = P.off + tail_begin;
        s.k1_n = r1 - tail_begin;
    }
    return s;
}

// descriptor ints of one phase (GRAM_PHASE_INTS) from a split, block indices relative to the stores
void write_phase_desc(const Split& s, const PhasePlan& P, int* d) {
    d[0] = (int)s.k0_row; d[1] = (int)s.k0_n; d[2] = (int)s.k1_row; d[3] = (int)s.k1_n;
    d[4] = (int)(s.c_lo - P.bmin[0]);  d[5] = (int)(s.c_hi - s.c_lo);
    d[6] = (int)(s.fa_lo - P.bmin[1]); d[7] = (int)(s.fa_hi - s.fa_lo);
    d[8] = (int)(s.fb_lo - P.bmin[1]); d[9] = (int)(s.fb_hi - s.fb_lo);
    if (d[5] == 0) d[4] = 0;
    if (d[7] == 0) d[6] = 0;
    if (d[9] == 0) d[8] = 0;
}

int upload_batch(bp_handle* h, const bp_window_batch* b, bool need_hf, Batch* out) {
    if (!h || !b) return fail(BP_ERR_INVALID, "null handle or batch");
    if (!h->has_market) return fail(BP_ERR_STATE, "no market uploaded: call bp_upload_market first");
    const int W = b->n_windows, n = b->rolling_window;
    if (W <= 0) return fail(BP_ERR_INVALID, "n_windows must be positive");
    if (n < 3) return fail(BP_ERR_INVALID, "rolling_window must be >= 3");
    if (!b->day_row || !b->span_days) return fail(BP_ERR_INVALID, "day_row / span_days missing");
    if (need_hf) {
        if (!b->hf_lo || !b->hf_hi) return fail(BP_ERR_INVALID, "hf_lo / hf_hi missing");
        if (h->R <= 0) return fail(BP_ERR_STATE, "the uploaded market has no intraday prices");
        if (b->resampled && !b->prior_n && !h->mcm_w) return fail(BP_ERR_STATE, "bp_set_resampled was given no MCM rows");
        if (!b->prior_n && (b->mcm_index < 0 || b->mcm_index >= h->n_mcm))
            return fail(BP_ERR_INVALID, "mcm_index %d out of range", b->mcm_index);
        if (b->mcm_rows < 0 || b->mcm_rows > n) return fail(BP_ERR_INVALID, "mcm_rows must be in [0, rolling_window]");
        if (b->prior_weights == 0 && !h->has_caps) return fail(BP_ERR_STATE, "value-weighted prior needs market caps");
    }
    const bool rs = b->resampled != 0;
    if (rs) {
        if (h->Rw <= 0) return fail(BP_ERR_STATE, "resampled windows need bp_set_resampled first");
        if (!b->extra_row || !b->caps_row) return fail(BP_ERR_INVALID, "resampled windows need extra_row and caps_row");
    }
    // ---- block grids for the window-overlap reuse of the Gram kernel
    const int npairs_t = ((h->N + GRAM_TILE - 1) / GRAM_TILE) * ((h->N + GRAM_TILE - 1) / GRAM_TILE + 1) / 2;
    PhasePlan plan[2];
    if (W >= h->reuse_min_windows) {
        const int K = n - 1;
        if (!rs) {          // resampled windows: shared weekly rows + one per-date row, no block reuse
            // Block sizes were tuned on the bench workload (N = 500): coarse/fine 128/32 for the 1007-row Jeffreys
            // window and 64/16 for the 251-row conjugate window gave 18.6 ms for the Gram stage, 512/64 + 128/32
            // 19.4 ms, a single level (128 / 64) 21.9 ms; neither the DMMA pipe (40-50 %) nor L2 (26-32 %) is
            // saturated then, the stage is bound by per-item latencies.
            if (K >= 512) { plan[1].blk[0] = 128; plan[1].blk[1] = 32; }
            else if (K >= 128) { plan[1].blk[0] = 64; plan[1].blk[1] = 16; }
        }
        if (need_hf) {
            // Day blocks: the HF look-back of every window covers whole days, so the sorted set of all hf_lo / hf_hi
            // values cuts the bars into blocks of which every window is a run -- whatever the bar calendar (half
            // days, holidays).  Worth it when the windows overlap (each block serves >= 2 windows on average).
            static const bool no_inner = getenv("BP_NO_INNER_BLOCKS") != nullptr;
            std::vector<int>& st = plan[0].starts;
            st.reserve(2 * (size_t)W);
            for (int w = 0; w < W; ++w) { st.push_back(b->hf_lo[w]); st.push_back(b->hf_hi[w]); }
            std::sort(st.begin(), st.end());
            st.erase(std::unique(st.begin(), st.end()), st.end());
            if (st.size() >= 3) {
                // stretches no window starts or ends in (the look-back of the first windows of a short batch) would be
                // single giant blocks contracted by one CTA per tile pair: cut them to the typical block length (a
                // window contains such a stretch entirely or not at all, so any cut inside it is a valid boundary)
                std::vector<int> gaps(st.size() - 1);
                for (size_t k = 0; k + 1 < st.size(); ++k) gaps[k] = st[k + 1] - st[k];
                std::nth_element(gaps.begin(), gaps.begin() + gaps.size() / 2, gaps.end());
                const int typical = std::max(8, gaps[gaps.size() / 2]);
                std::vector<int> cut;
                cut.reserve(st.size());
                for (size_t k = 0; k + 1 < st.size(); ++k) {
                    cut.push_back(st[k]);
                    const int gap = st[k + 1] - st[k];
                    if (gap > 2 * typical) {
                        const int pieces = gap / typical;
                        for (int q = 1; q < pieces; ++q) cut.push_back(st[k] + (int)((long long)gap * q / pieces));
                    }
                }
                cut.push_back(st.back());
                st.swap(cut);
            }
            const long long nb = (long long)st.size() - 1;
            long long sumD = 0;
            int minD = 1 << 30, maxD = 0, min_rows = 1 << 30;
            for (int w = 0; w < W; ++w) {
                const int D = (int)(std::lower_bound(st.begin(), st.end(), b->hf_hi[w]) - std::lower_bound(st.begin(), st.end(), b->hf_lo[w]));
                sumD += D; minD = std::min(minD, D); maxD = std::max(maxD, D);
            }
            for (long long k = 0; k < nb; ++k) min_rows = std::min(min_rows, st[k + 1] - st[k]);
            const bool day_blocks = !no_inner && nb >= 2 && minD >= 2 && min_rows >= 8 && sumD >= 2 * nb && nb <= h->hf_extra_rows;
            if (day_blocks) {
                plan[0].inner = true;
                plan[0].blk[0] = 1;                 // level 0 in use (block sizes vary: see starts)
                const int minL = minD - 1, maxL = maxD - 1;       // full days (overnight return included) per window
                static const bool no_presum = getenv("BP_NO_PRESUM") != nullptr;
                if (!no_presum && h->presum_min_days > 0 && minD >= h->presum_min_days && maxL <= 2 * minL &&
                    3 * (size_t)nb * npairs_t * GRAM_BLOCK_TILE_DOUBLES * sizeof(double) <= h->ws_limit)
                    plan[0].chunk = minL;
            } else {
                st.clear();
                const int H0 = b->hf_hi[0] - b->hf_lo[0];
                if (H0 >= 3 * 64) {
                    plan[0].blk[0] = 64;
                    plan[0].blk[1] = 16;
                }
            }
        }
    }
    auto phase_rows = [&](int ph, int w, int& row0, int& rows) {
        row0 = ph == 0 ? b->hf_lo[w] + 1 : b->day_row[w] - n + 2;
        rows = ph == 0 ? b->hf_hi[w] - b->hf_lo[w] - 1 : n - 1;
    };
    // split of window w's rows in phase ph; inner-block plans take every day of the window as a whole block
    auto split_of = [&](int ph, int w) {
        int row0, rows;
        phase_rows(ph, w, row0, rows);
        if (ph == 0 && plan[0].inner) {
            const std::vector<int>& st = plan[0].starts;
            Split sp;
            sp.c_lo = std::lower_bound(st.begin(), st.end(), b->hf_lo[w]) - st.begin();
            sp.c_hi = std::lower_bound(st.begin(), st.end(), b->hf_hi[w]) - st.begin();
            return sp;
        }
        return split_rows(row0, rows, plan[ph]);
    };
    // phase B (daily rows): the run of whole blocks inside a window is the same for many consecutive windows, so
    // each distinct run [lo, hi) is summed once (range_sum_kernel) and a window adds one tile per run
    std::map<std::pair<long long, long long>, int> range_id[2];
    // ... and the (coarse, fine, fine) run triple of a window is combined once per distinct triple (combine_runs_kernel)
    typedef std::array<long long, 6> RunKey;
    std::map<RunKey, int> full_id;
    // first pass: block ranges touched by the windows, per level
    for (int ph = 0; ph < 2; ++ph) {
        if (plan[ph].blk[0] <= 0 && plan[ph].blk[1] <= 0) continue;
        if (ph == 0 && !need_hf) { plan[ph] = PhasePlan(); continue; }
        long long lo[2] = {1LL << 40, 1LL << 40}, hi[2] = {-(1LL << 40), -(1LL << 40)};
        for (int w = 0; w < W; ++w) {
            if (w % h->work_stride != 0) continue;      // chain path: only the base windows go through the Gram kernel
            const Split sp = split_of(ph, w);
            if (sp.c_hi > sp.c_lo) { lo[0] = std::min(lo[0], sp.c_lo); hi[0] = std::max(hi[0], sp.c_hi); }
            if (sp.fa_hi > sp.fa_lo) { lo[1] = std::min(lo[1], sp.fa_lo); hi[1] = std::max(hi[1], sp.fa_hi); }
            if (sp.fb_hi > sp.fb_lo) { lo[1] = std::min(lo[1], sp.fb_lo); hi[1] = std::max(hi[1], sp.fb_hi); }
            if (ph == 1) {
                if (sp.c_hi > sp.c_lo) range_id[0].emplace(std::make_pair(sp.c_lo, sp.c_hi), 0);
                if (sp.fa_hi > sp.fa_lo) range_id[1].emplace(std::make_pair(sp.fa_lo, sp.fa_hi), 0);
                if (sp.fb_hi > sp.fb_lo) range_id[1].emplace(std::make_pair(sp.fb_lo, sp.fb_hi), 0);
                if (sp.c_hi > sp.c_lo || sp.fa_hi > sp.fa_lo || sp.fb_hi > sp.fb_lo)
                    full_id.emplace(RunKey{sp.c_lo, sp.c_hi, sp.fa_lo, sp.fa_hi, sp.fb_lo, sp.fb_hi}, 0);
            }
        }
        for (int l = 0; l < 2; ++l) {
            if (hi[l] <= lo[l]) { plan[ph].blk[l] = 0; plan[ph].nb[l] = 0; continue; }
            plan[ph].bmin[l] = (int)lo[l];
            plan[ph].nb[l] = (int)(hi[l] - lo[l]);
            size_t need = (size_t)plan[ph].nb[l] * npairs_t * GRAM_BLOCK_TILE_DOUBLES;
            if (need * sizeof(double) > h->ws_limit / 2) {      // too big: give up the reuse of this phase
                plan[ph] = PhasePlan();
                break;
            }
            if (ph == 0 && l == 0 && plan[0].chunk > 0) need *= 3;      // + suffix' and prefix tiles of every block
            if (need > h->store_cap[ph][l]) {
                CU_TRY(cudaStreamSynchronize(h->stream));
                cudaFree(h->store[ph][l]);
                h->store[ph][l] = nullptr;
                h->store_cap[ph][l] = 0;
                CU_TRY(cudaMalloc(&h->store[ph][l], need * sizeof(double)));
                h->store_cap[ph][l] = need;
            }
        }
    }
    if (plan[0].inner && plan[0].nb[0] != (int)plan[0].starts.size() - 1)
        return fail(BP_ERR_STATE, "internal: day-block plan out of step with its boundaries");
    const bool presum = plan[0].inner && plan[0].chunk > 0;
    const int nb0 = plan[0].inner ? plan[0].nb[0] : 0;
    if (presum) {
        const size_t need = 3 * (size_t)nb0 * h->ld;
        if (need > h->vstore_cap) {
            CU_TRY(cudaStreamSynchronize(h->stream));
            cudaFree(h->vstore);
            h->vstore = nullptr;
            h->vstore_cap = 0;
            CU_TRY(cudaMalloc(&h->vstore, need * sizeof(double)));
            h->vstore_cap = need;
        }
    }
    int nranges[2] = {0, 0};
    for (int l = 0; l < 2; ++l) {
        if (plan[1].blk[l] <= 0) { range_id[l].clear(); continue; }
        for (auto& kv : range_id[l]) kv.second = nranges[l]++;
        const size_t need = (size_t)nranges[l] * npairs_t * GRAM_BLOCK_TILE_DOUBLES;
        if (need > h->rstore_cap[l]) {
            CU_TRY(cudaStreamSynchronize(h->stream));
            cudaFree(h->rstore[l]);
            h->rstore[l] = nullptr;
            h->rstore_cap[l] = 0;
            CU_TRY(cudaMalloc(&h->rstore[l], need * sizeof(double)));
            h->rstore_cap[l] = need;
        }
    }
    int nfull = 0;
    if (plan[1].blk[0] <= 0 && plan[1].blk[1] <= 0) full_id.clear();
    for (auto& kv : full_id) kv.second = nfull++;
    {
        const size_t need = (size_t)nfull * npairs_t * GRAM_BLOCK_TILE_DOUBLES;
        if (need > h->fstore_cap) {
            CU_TRY(cudaStreamSynchronize(h->stream));
            cudaFree(h->fstore);
            h->fstore = nullptr;
            h->fstore_cap = 0;
            CU_TRY(cudaMalloc(&h->fstore, need * sizeof(double)));
            h->fstore_cap = need;
        }
    }
    const int nb_total = plan[0].nb[0] + plan[0].nb[1] + plan[1].nb[0] + plan[1].nb[1];
    const size_t ints_needed = (size_t)(7 + GRAM_DESC_INTS) * W + (size_t)GRAM_DESC_INTS * nb_total +
                               2 * (size_t)(nranges[0] + nranges[1]) + 3 * (size_t)nfull +
                               (plan[0].inner ? (size_t)nb0 + 1 : 0) + (presum ? 3 * (size_t)W : 0);
    const int slot = h->desc_turn ^= 1;
    if (h->ev_desc_armed[slot]) {
        CU_TRY(cudaEventSynchronize(h->ev_desc[slot]));      // its staging buffer has been fetched
        h->ev_desc_armed[slot] = false;
    }
    if (ints_needed > h->desc_cap_slot[slot]) {
        if (h->desc_slot[slot]) {
            CU_TRY(cudaStreamSynchronize(h->stream));
            cudaFree(h->desc_slot[slot]);
            cudaFreeHost(h->desc_host_slot[slot]);
        }
        h->desc_slot[slot] = nullptr;
        h->desc_host_slot[slot] = nullptr;
        h->desc_cap_slot[slot] = 0;
        CU_TRY(cudaMalloc(&h->desc_slot[slot], sizeof(int) * ints_needed));
        CU_TRY(cudaHostAlloc(&h->desc_host_slot[slot], sizeof(int) * ints_needed, cudaHostAllocDefault));
        h->desc_cap_slot[slot] = ints_needed;
    }
    h->desc = h->desc_slot[slot];
    h->desc_host = h->desc_host_slot[slot];
    h->desc_cap = h->desc_cap_slot[slot];
    int* host = h->desc_host;
    memset(host, 0, sizeof(int) * ints_needed);
    int* gd = host + (size_t)7 * W;                  // Gram job descriptors
    int max_m = 0;
    for (int w = 0; w < W; ++w) {
        const int dr = b->day_row[w];
        if (rs) {
            // shared weekly rows dr-(n-2)+1 .. dr (row 0 is the zero row of the first week) + one per-date row
            if (dr - (n - 2) + 1 < 1 || dr >= h->Rw)
                return fail(BP_ERR_INVALID, "window %d: resampled day_row %d needs %d prior rows inside [1,%d)", w, dr, n - 2, h->Rw);
            if (b->extra_row[w] < 0 || b->extra_row[w] >= h->Rw || b->caps_row[w] < 0 || b->caps_row[w] >= h->D)
                return fail(BP_ERR_INVALID, "window %d: extra_row / caps_row out of range", w);
            host[(size_t)5 * W + w] = b->extra_row[w];
            host[(size_t)6 * W + w] = b->caps_row[w];
        } else {
            if (dr < n - 1 || dr >= h->D)
                return fail(BP_ERR_INVALID, "window %d: day_row %d needs %d prior price rows inside [0,%d)", w, dr, n - 1, h->D);
            host[(size_t)5 * W + w] = -1;
            host[(size_t)6 * W + w] = dr;
        }
        if (need_hf && !b->prior_n && dr < (b->mcm_rows ? b->mcm_rows : n) - 1 - (rs ? 1 : 0))
            return fail(BP_ERR_INVALID, "window %d: not enough MCM observations before day_row %d", w, dr);
        if (b->span_days[w] <= 0) return fail(BP_ERR_INVALID, "window %d: span_days must be positive", w);
        host[w] = dr;
        host[(size_t)W + w] = b->span_days[w];
        host[(size_t)2 * W + w] = rs ? dr - (n - 2) + 1 : dr - n + 2;   // first RETURN row (F2: n prices -> n-1 returns)
        if (need_hf) {
            const int lo = b->hf_lo[w], hi = b->hf_hi[w];
            const int m = hi - lo - 1;
            if (lo < 0 || (long long)hi > h->R || m < 2)
                return fail(BP_ERR_INVALID, "window %d: intraday rows [%d,%d) give %d returns (need >= 2, rows < %lld)", w, lo, hi, m, h->R);
            host[(size_t)3 * W + w] = lo + 1;        // first HF return row: the window's first bar has no return (F5)
            host[(size_t)4 * W + w] = m;
            max_m = std::max(max_m, m);
            int* dA = gd + (size_t)w * GRAM_DESC_INTS;
            write_phase_desc(split_of(0, w), plan[0], dA);
            if (presum) {
                // suffix'[first day] (+ one whole chunk) (+ prefix[last day]): explicit tile ids into the unified store
                // [inner | suffix' | prefix]; the same ids address the scanned column sums (hf_vids)
                const int C = plan[0].chunk, c_lo = dA[4], last = dA[4] + dA[5] - 1;
                const int qa = c_lo / C, s_next = std::min((qa + 1) * C, nb0);
                int ids[3] = {nb0 + c_lo, -1, -1};
                if (s_next <= last) {
                    const int ql = last / C;
                    if (ql > qa + 2) return fail(BP_ERR_STATE, "internal: pre-summed run spans more than three chunks");
                    if (ql == qa + 2) ids[1] = 2 * nb0 + (qa + 2) * C - 1;
                    ids[2] = 2 * nb0 + last;
                }
                dA[0] = dA[1] = dA[2] = dA[3] = 0;
                for (int k = 0; k < 3; ++k) { dA[4 + 2 * k] = ids[k] >= 0 ? ids[k] : 0; dA[5 + 2 * k] = ids[k] >= 0 ? 1 : 0; }
                int* vid = host + ((size_t)(7 + GRAM_DESC_INTS) * W + (size_t)GRAM_DESC_INTS * nb_total +
                                   2 * (size_t)(nranges[0] + nranges[1]) + 3 * (size_t)nfull + (size_t)nb0 + 1) + 3 * (size_t)w;
                vid[0] = ids[0]; vid[1] = ids[1]; vid[2] = ids[2];
            } else if (plan[0].inner) {
                // overnight returns of days 2..last of the window: a run of dA[5]-1 gathered rows behind the matrix
                dA[0] = (int)(h->R + dA[4] + 1);
                dA[1] = dA[5] - 1;
            }
        }
        if (rs) {
            int* dB = gd + (size_t)w * GRAM_DESC_INTS + GRAM_PHASE_INTS;
            dB[0] = dr - (n - 2) + 1; dB[1] = n - 2;       // shared weekly rows
            dB[2] = b->extra_row[w];  dB[3] = 1;           // the trade date's own row
        } else if (w % h->work_stride == 0) {
            const Split sp = split_rows(dr - n + 2, n - 1, plan[1]);
            int* dB = gd + (size_t)w * GRAM_DESC_INTS + GRAM_PHASE_INTS;
            write_phase_desc(sp, plan[1], dB);
            // one pre-summed tile per run of whole blocks
            // ... combined into one tile per distinct (coarse, fine, fine) triple: the coarse slot points into fstore
            if (dB[5] > 0 || dB[7] > 0 || dB[9] > 0) {
                dB[4] = full_id[RunKey{sp.c_lo, sp.c_hi, sp.fa_lo, sp.fa_hi, sp.fb_lo, sp.fb_hi}];
                dB[5] = 1;
                dB[6] = dB[7] = dB[8] = dB[9] = 0;
            }
        }
        const int* d = gd + (size_t)w * GRAM_DESC_INTS;
        auto r8 = [](int r) { return (r + 7) / 8 * 8; };
        for (int ph = 0; ph < 2 && w % h->work_stride == 0; ++ph) {
            const int* dp = d + ph * GRAM_PHASE_INTS;
            h->work_k_rows += r8(dp[1]) + r8(dp[3]);
            h->work_add_blocks += dp[5] + dp[7] + dp[9];
        }
        h->work_full_rows += (need_hf ? b->hf_hi[w] - b->hf_lo[w] - 1 : 0) + (n - 1);
    }
    for (int ph = 0; ph < 2; ++ph)
        for (int l = 0; l < 2; ++l) {
            if (ph == 0 && l == 0 && plan[0].inner) {
                for (int k = 0; k < nb0; ++k) h->work_pre_rows += (plan[0].starts[k + 1] - plan[0].starts[k] - 1 + 7) / 8 * 8;
                continue;
            }
            h->work_pre_rows += (double)plan[ph].nb[l] * ((plan[ph].blk[l] + 7) / 8 * 8);
        }
    if (presum) h->work_scan_tiles += (double)nb0;
    // descriptors of the block precompute launches: one pseudo-window per block, rows of that block only
    int* bd = gd + (size_t)W * GRAM_DESC_INTS;
    for (int ph = 0; ph < 2; ++ph)
        for (int l = 0; l < 2; ++l) {
            for (int k = 0; k < plan[ph].nb[l]; ++k) {
                int* d = bd + (size_t)k * GRAM_DESC_INTS + GRAM_PHASE_INTS * ph;
                if (plan[ph].inner) {           // the day without its overnight return
                    d[0] = plan[ph].starts[k] + 1;
                    d[1] = plan[ph].starts[k + 1] - plan[ph].starts[k] - 1;
                } else {
                    d[0] = plan[ph].off + (plan[ph].bmin[l] + k) * plan[ph].blk[l];
                    d[1] = plan[ph].blk[l];
                }
            }
            out->bdesc[ph][l] = plan[ph].nb[l] ? h->desc + (bd - host) : nullptr;
            out->nblocks[ph][l] = plan[ph].nb[l];
            bd += (size_t)plan[ph].nb[l] * GRAM_DESC_INTS;
        }
    // block runs to pre-sum (store-relative block indices)
    for (int l = 0; l < 2; ++l) {
        for (const auto& kv : range_id[l]) {
            bd[2 * kv.second] = (int)(kv.first.first - plan[1].bmin[l]);
            bd[2 * kv.second + 1] = (int)(kv.first.second - plan[1].bmin[l]);
        }
        out->rdesc[l] = nranges[l] ? h->desc + (bd - host) : nullptr;
        out->nranges[l] = nranges[l];
        bd += 2 * (size_t)nranges[l];
    }
    for (const auto& kv : full_id) {
        const RunKey& k = kv.first;
        int* d = bd + 3 * (size_t)kv.second;
        d[0] = k[1] > k[0] ? range_id[0][std::make_pair(k[0], k[1])] : -1;
        d[1] = k[3] > k[2] ? range_id[1][std::make_pair(k[2], k[3])] : -1;
        d[2] = k[5] > k[4] ? range_id[1][std::make_pair(k[4], k[5])] : -1;
    }
    out->fdesc = nfull ? h->desc + (bd - host) : nullptr;
    out->nfull = nfull;
    bd += 3 * (size_t)nfull;
    if (plan[0].inner) {
        for (int k = 0; k <= nb0; ++k) bd[k] = plan[0].starts[k];
        out->hf_starts_dev = h->desc + (bd - host);
        out->hf_starts = plan[0].starts;
        bd += (size_t)nb0 + 1;
        if (presum) {
            out->hf_vids = h->desc + (bd - host);       // filled per window above
            bd += 3 * (size_t)W;
        }
    }
    out->presum = presum;
    out->presum_chunk = plan[0].chunk;
    out->gdesc = h->desc + (size_t)7 * W;
    {
        // banded-GEMM daily pass: worth it when 32 consecutive windows share almost all of their rows
        static const bool no_band = getenv("BP_NO_BAND") != nullptr;
        bool ok = !no_band && !rs && W >= 32 && n >= 16;
        for (int w = 1; ok && w < W; ++w) ok = b->day_row[w] >= b->day_row[w - 1];
        if (ok) ok = (long long)b->day_row[W - 1] - b->day_row[0] <= 2LL * W;
        out->band_ok = ok;
    }
    out->hf_off = plan[0].off;
    for (int l = 0; l < 2; ++l) { out->hf_bmin[l] = plan[0].bmin[l]; out->hf_blk[l] = plan[0].blk[l]; }
    out->hf_inner = plan[0].inner;
    out->hf_nb0 = plan[0].nb[0];
    out->resampled = rs;
    out->extra_row = rs ? h->desc + 5 * (size_t)W : nullptr;
    out->caps_row = h->desc + 6 * (size_t)W;
    // zero-copy fetch by a kernel on the compute stream (not the copy engine, see fetch_ints_kernel); the
    // staging buffer of this slot is rewritten two batches later, after this event
    launch_fetch_ints(host, h->desc, (long long)ints_needed, h->stream);
    h->launches++;
    CU_TRY(cudaGetLastError());
    if (!h->ev_desc[slot]) CU_TRY(cudaEventCreateWithFlags(&h->ev_desc[slot], cudaEventDisableTiming));
    CU_TRY(cudaEventRecord(h->ev_desc[slot], h->stream));
    h->ev_desc_armed[slot] = true;
    if (need_hf && b->prior_n) {
        if (W > h->prior_n_cap) {
            cudaFree(h->prior_n);
            h->prior_n = nullptr;
            h->prior_n_cap = 0;
            CU_TRY(cudaMalloc(&h->prior_n, sizeof(double) * (size_t)W));
            h->prior_n_cap = W;
        }
        CU_TRY(cudaMemcpyAsync(h->prior_n, b->prior_n, sizeof(double) * (size_t)W, cudaMemcpyHostToDevice, h->stream));
        CU_TRY(cudaStreamSynchronize(h->stream));
        out->prior_n = h->prior_n;
    }
    out->mcm_rows = b->mcm_rows ? b->mcm_rows : n;
    out->W = W;
    out->n = n;
    out->max_m = max_m;
    out->has_hf = need_hf;
    out->day_row = h->desc;
    out->span = h->desc + W;
    out->row0 = h->desc + 2 * (size_t)W;
    out->hf_row0 = h->desc + 3 * (size_t)W;
    out->hf_m = h->desc + 4 * (size_t)W;
    return BP_OK;
}

// dense [Wc][N] copy of a strided device vector to a host or device destination
int emit_vec(bp_handle* h, const double* src, const Layout& L, int Wc, double* dst) {
    if (!dst) return BP_OK;
    if (is_device_ptr(dst)) {
        launch_unpack_vec(src, L.ldv, L.N, Wc, dst, h->stream);
        h->launches++;
    } else {
        const size_t bytes = sizeof(double) * (size_t)Wc * L.N;
        // the single staging buffer is reused: drain the previous D2H copy before overwriting it
        if (h->need_sync) CU_TRY(cudaStreamSynchronize(h->stream));
        int rc = ensure_stage(h, bytes);
        if (rc) return rc;
        launch_unpack_vec(src, L.ldv, L.N, Wc, reinterpret_cast<double*>(h->stage), h->stream);
        h->launches++;
        CU_TRY(cudaMemcpyAsync(dst, h->stage, bytes, cudaMemcpyDeviceToHost, h->stream));
        h->need_sync = true;
    }
    return BP_OK;
}

int emit_sym(bp_handle* h, const double* S, const Layout& L, int Wc, double* dst) {
    if (!dst) return BP_OK;
    if (is_device_ptr(dst)) {
        launch_unpack_sym(S, L.win_stride, L.ldS, L.N, Wc, dst, h->stream);
        h->launches++;
    } else {
        const size_t bytes = sizeof(double) * (size_t)Wc * L.N * L.N;
        if (h->need_sync) CU_TRY(cudaStreamSynchronize(h->stream));
        int rc = ensure_stage(h, bytes);
        if (rc) return rc;
        launch_unpack_sym(S, L.win_stride, L.ldS, L.N, Wc, reinterpret_cast<double*>(h->stage), h->stream);
        h->launches++;
        CU_TRY(cudaMemcpyAsync(dst, h->stage, bytes, cudaMemcpyDeviceToHost, h->stream));
        h->need_sync = true;
    }
    return BP_OK;
}

int emit_raw(bp_handle* h, const void* src, size_t bytes, void* dst) {
    if (!dst) return BP_OK;
    const bool dev = is_device_ptr(dst);
    CU_TRY(cudaMemcpyAsync(dst, src, bytes, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->stream));
    if (!dev) h->need_sync = true;
    return BP_OK;
}

PrepParams prep_params(const bp_handle* h, const bp_window_batch* b, const Batch& B, const Layout& L, const Chunk& c,
                       int w0, int mode) {
    PrepParams p{};
    p.mode = mode;
    p.n_assets = h->N;
    p.n_window = B.n;
    p.ld = h->ld;
    p.ldv = L.ldv;
    p.prior_kind = b->prior_weights == 0 ? BP_PRIOR_VW : BP_PRIOR_EW;
    p.mcm_scaling = b->mcm_scaling;
    p.lr_daily = B.resampled ? h->lr_w : h->lr_d;
    p.lr_hf = h->lr_hf;
    p.caps = h->has_caps ? h->caps : nullptr;
    p.ld_caps = h->N;
    p.mcm = (mode == BP_MODE_CONJUGATE && h->mcm) ? h->mcm + (size_t)b->mcm_index * h->D : nullptr;
    if (B.resampled) p.mcm = (mode == BP_MODE_CONJUGATE && h->mcm_w) ? h->mcm_w + (size_t)b->mcm_index * h->Rw : nullptr;
    p.mcm_rows = B.mcm_rows;
    p.prior_n = B.prior_n ? B.prior_n + w0 : nullptr;
    p.rf_row = B.resampled ? h->rf_w : h->rf_row;
    p.day_row = B.day_row + w0;
    p.extra_row = B.extra_row ? B.extra_row + w0 : nullptr;
    p.caps_row = B.caps_row ? B.caps_row + w0 : nullptr;
    p.span_days = B.span + w0;
    p.hf_row0 = B.hf_row0 + w0;
    p.hf_m = B.hf_m + w0;
    p.t = c.t; p.pvec = c.pvec; p.gvec = c.gvec; p.rhs = c.rhs; p.w0 = c.w0; p.s0w0 = c.s0w0;
    p.scal = c.scal;
    p.y_ws = c.y;
    p.y_stride = L.y_stride;
    p.hf_presum = (mode == BP_MODE_CONJUGATE && B.presum) ? 1 : 0;
    p.hf_vsum = h->vstore;
    p.hf_vids = B.hf_vids ? B.hf_vids + 3 * (size_t)w0 : nullptr;
    p.use_band = B.band_ok ? 1 : 0;
    p.band_ld = round_up(B.n - 1, 2);
    p.band_aw = B.band_ok ? h->band_aw + (size_t)w0 * p.band_ld : nullptr;
    p.band_stats = B.band_ok ? h->band_stats + 2 * (size_t)w0 : nullptr;
    return p;
}

enum GramKind { GRAM_T, GRAM_S0, GRAM_S1, GRAM_J };

GramParams gram_params(const bp_handle* h, const Batch& B, const Layout& L, const Chunk& c, int w0, int Wc, GramKind kind) {
    GramParams g{};
    g.n_windows = Wc;
    g.n_assets = h->N;
    g.ldS = L.ldS;
    g.win_stride = L.win_stride;
    g.ldv = L.ldv;
    g.mirror = 0;
    g.scal = c.scal;
    g.out = c.S;
    g.desc = B.gdesc + (size_t)w0 * GRAM_DESC_INTS;
    for (int l = 0; l < 2; ++l) {
        g.store[0][l] = h->store[0][l];          // intraday: one tile per whole block
        g.store[1][l] = l == 0 ? h->fstore : h->rstore[l];      // daily: ONE combined tile per window (runs of whole blocks, pre-summed)
    }
    const bool hf = kind == GRAM_S0 || kind == GRAM_S1;
    const bool daily = kind != GRAM_S0;
    if (B.presum) {
        g.store[0][1] = h->store[0][0];          // explicit tile ids into the unified [inner | suffix' | prefix] store
        if (hf) {
            g.mv_part = c.mv;
            g.mv_w0 = c.w0;
        }
    }
    g.use_phaseA = hf;
    g.use_phaseB = daily;
    if (hf) {
        g.use_alpha = 1;
        g.use_beta = 1;          // beta = alpha*m, g = hbar
        g.gvec = c.gvec;
    }
    if (daily) g.pvec = c.pvec;
    if (kind == GRAM_J) {
        g.use_beta = 1;          // beta = 1/n, g = t
        g.gvec = c.gvec;
    }
    return g;
}

// block precompute: the Gram tile of the whole blocks [k0[l], k1[l]) of a phase / level (default: all of them),
// written fragment-major into its store
int run_block_precompute(bp_handle* h, const Batch& B, int ph, const int* k0 = nullptr, const int* k1 = nullptr) {
    const int nt = (h->N + GRAM_TILE - 1) / GRAM_TILE;
    const size_t tile_doubles = (size_t)(nt * (nt + 1) / 2) * GRAM_BLOCK_TILE_DOUBLES;
    for (int l = 0; l < 2; ++l) {
        const int b0 = k0 ? k0[l] : 0, b1 = k1 ? k1[l] : B.nblocks[ph][l];
        if (b1 <= b0) continue;
        GramParams g{};
        g.n_windows = b1 - b0;
        g.n_assets = h->N;
        g.desc = B.bdesc[ph][l] + (size_t)b0 * GRAM_DESC_INTS;
        g.use_phaseA = ph == 0;
        g.use_phaseB = ph == 1;
        g.tile_store_out = 1;
        g.out = h->store[ph][l] + (size_t)b0 * tile_doubles;
        StageTimer tm(h, BP_STAGE_GRAM);
        CU_TRY(launch_gram(g, h->map_hf, h->map_d, h->sm_count, h->stream));
        h->launches++;
        if (ph == 1 && B.nranges[l] > 0) {
            launch_range_sum(h->store[1][l], B.rdesc[l], B.nranges[l], nt * (nt + 1) / 2, h->rstore[l], h->stream);
            h->launches++;
            CU_TRY(cudaGetLastError());
        }
    }
    if (ph == 1 && !k0 && B.nfull > 0) {
        StageTimer tm(h, BP_STAGE_GRAM);
        launch_combine_runs(h->rstore[0], h->rstore[1], B.fdesc, B.nfull, nt * (nt + 1) / 2, h->fstore, h->stream);
        h->launches++;
        CU_TRY(cudaGetLastError());
    }
    return BP_OK;
}

int run_gram(bp_handle* h, const GramParams& g, bool resampled) {
    StageTimer tm(h, BP_STAGE_GRAM);
    CU_TRY(launch_gram(g, h->map_hf, resampled ? h->map_w : h->map_d, h->sm_count, h->stream));
    h->launches++;
    return BP_OK;
}

int finish(bp_handle* h) {
    if (h->need_sync && !h->async_outputs) {
        CU_TRY(cudaStreamSynchronize(h->stream));
        h->need_sync = false;
    }
    return BP_OK;
}

// shared driver of bp_conjugate_batched / bp_jeffreys_batched / bp_stats_batched / bp_hf_cov_batched
int run_batches(bp_handle* h, const bp_window_batch* b, const bp_outputs* out, int mode, bool solve, int estimator = BP_EST_NONE) {
    Batch B;
    if (!h || !b) return fail(BP_ERR_INVALID, "null handle or batch");
    // every allocation / launch below (upload_batch included) must land on the handle's device, whatever the
    // calling thread's current device is (two handles on two GPUs in one thread, or torch having switched)
    CU_TRY(cudaSetDevice(h->device));
    // Jeffreys batches of consecutive trade dates: factorise every G-th window only (jeffreys_chain.cu)
    int G = 0;
    if (h && b && out && mode == BP_MODE_JEFFREYS && (estimator == BP_EST_NONE || estimator == BP_EST_JORION) && solve && !out->T && !out->S0 && !out->S1 &&
        !b->resampled && h->chain_group >= 2 && b->n_windows >= 2 * h->chain_group && b->day_row && h->has_market &&
        chain_smem_bytes(h->N) <= (size_t)227 * 1024) {
        bool consecutive = true;
        for (int w = 1; consecutive && w < b->n_windows; ++w) consecutive = b->day_row[w] == b->day_row[0] + w;
        const Layout L0 = make_layout(h, 0);
        if (consecutive && (size_t)b->n_windows * L0.per_window <= h->ws_limit) G = std::min(h->chain_group, chain_max_group());
        // Conditioning guard: window b+k is the base with k rows removed and k rows added; the down-dated matrix
        // has rank n-1-k, and its Woodbury pivots -1 + l'J_b^-1 l tend to 0 as n-1-k approaches N (leverage 1),
        // where the small system loses digits while no pivot is exactly zero.  Chain only with a margin of rows;
        // closer to the rank limit every window is factorised on its own (the path the parity tests pin).
        if (G >= 2 && b->rolling_window - 1 < h->N + 2 * G + 8) G = 0;
    }
    if (h) h->work_stride = G >= 2 ? G : 1;
    int rc = upload_batch(h, b, mode == BP_MODE_CONJUGATE, &B);
    if (h) h->work_stride = 1;
    if (rc) return rc;
    const Layout L = make_layout(h, B.presum ? 0 : B.max_m, B.presum);
    int Wc = (int)std::min<size_t>((size_t)B.W, std::max<size_t>(1, h->ws_limit / L.per_window));
    // (the descriptors were planned for the base windows only: the chain path must not be abandoned after upload_batch;
    // the check above uses the same layout, so this cannot trigger unless the two are changed apart)
    if (Wc < B.W && G >= 2) return fail(BP_ERR_STATE, "internal: chained batch does not fit one workspace chunk");
    // Pipelined against a segmented asynchronous intraday upload: prep + Gram of the windows whose bars have
    // arrived run while the rest is still being copied; the solve follows for all windows at once.
    const bool pipelined = mode == BP_MODE_CONJUGATE && h->hf_pending && h->n_seg > 1 && h->seg_waited < h->n_seg &&
                           solve && !out->T && !out->S0 && Wc >= B.W && !B.presum;
    if (mode == BP_MODE_CONJUGATE && !pipelined) {
        long long need = 0;                   // intraday price rows this batch reads
        for (int w = 0; w < B.W; ++w) need = std::max<long long>(need, b->hf_hi[w]);
        if ((rc = wait_hf_rows(h, need))) return rc;
    }
    rc = ensure_ws(h, (size_t)Wc * L.per_window + 256);
    if (rc) return rc;
    const Chunk c = carve(h->ws, L, Wc);
    const int N = h->N;
    if (B.band_ok) {
        const size_t need_aw = (size_t)B.W * round_up(B.n - 1, 2), need_st = 2 * (size_t)B.W;
        if (need_aw > h->band_aw_cap || need_st > h->band_stats_cap) {
            CU_TRY(cudaStreamSynchronize(h->stream));
            cudaFree(h->band_aw);
            cudaFree(h->band_stats);
            h->band_aw = h->band_stats = nullptr;
            h->band_aw_cap = h->band_stats_cap = 0;
            CU_TRY(cudaMalloc(&h->band_aw, sizeof(double) * need_aw));
            CU_TRY(cudaMalloc(&h->band_stats, sizeof(double) * need_st));
            h->band_aw_cap = need_aw;
            h->band_stats_cap = need_st;
        }
    }
    if (solve && (mode == BP_MODE_CONJUGATE || estimator != BP_EST_NONE) && !(b->risk_aversion != 0.0))
        return fail(BP_ERR_INVALID, "risk_aversion must be non-zero");
    if (estimator == BP_EST_JORION && B.n - 1 - N - 2 <= 0)
        return fail(BP_ERR_INVALID, "Jorion needs T - N - 2 > 0 (T = rolling_window - 1 returns, :879)");

    const bool any_gram = out->T || out->S0 || out->S1 || solve;
    if (any_gram && !pipelined) {
        if (mode == BP_MODE_CONJUGATE && B.hf_inner && !B.presum) {
            launch_gather_rows_indexed(h->lr_hf, h->ld, B.hf_starts_dev, h->R, 0, B.hf_nb0, h->stream);
            h->launches++;
            CU_TRY(cudaGetLastError());
        }
        if (mode == BP_MODE_CONJUGATE && (out->S0 || out->S1 || solve || B.presum) && (rc = run_block_precompute(h, B, 0))) return rc;
        if (mode == BP_MODE_CONJUGATE && B.presum) {
            // scans of the day blocks: column sums (vectors) and Gram tiles, suffix' and prefix per chunk
            StageTimer tm(h, BP_STAGE_GRAM);
            const int nt = (N + GRAM_TILE - 1) / GRAM_TILE;
            launch_block_col_sums(h->lr_hf, h->ld, B.hf_starts_dev, 0, B.hf_nb0, h->vstore, h->stream);
            launch_vec_scan(h->vstore, h->ld, h->lr_hf, B.hf_starts_dev, B.hf_nb0, B.presum_chunk, h->stream);
            launch_tile_scan(h->store[0][0], nt * (nt + 1) / 2, nt, h->lr_hf, h->ld, B.hf_starts_dev, B.hf_nb0, B.presum_chunk, h->stream);
            h->launches += 3;
            CU_TRY(cudaGetLastError());
        }
        if ((out->T || out->S1 || solve) && (rc = run_block_precompute(h, B, 1))) return rc;
    }
    // Cholesky + solves of the windows [ws, ws+wn) of the workspace
    auto solve_range = [&](int ws, int wn, int stride = 1) -> int {
        if (wn <= 0) return BP_OK;
        const Chunk cs = chunk_at(c, L, ws % Wc);
        SolveParams sp{};
        sp.n_windows = wn;
        sp.w_stride = stride;
        sp.n_assets = N;
        sp.ldS = L.ldS;
        sp.win_stride = L.win_stride;
        sp.ldv = L.ldv;
        sp.mode = mode;
        sp.estimator = estimator;
        sp.n_returns = B.n - 1;
        sp.inv_gamma = 1.0 / b->risk_aversion;         // the reference evaluates (1/gamma) * nu (:836,:849)
        sp.S = cs.S;
        sp.rhs = cs.rhs;
        sp.scal = cs.scal;
        sp.w1 = cs.w1;
        sp.nu = cs.nu;
        sp.weights = cs.weights;
        sp.status = cs.status;
        CUtensorMap smap;
        int rcs = make_solve_map(h, &smap, cs.S, ((long long)(wn - 1) * stride + 1) * L.rowsS, L.ldS);
        if (rcs) return rcs;
        {
            StageTimer tm(h, BP_STAGE_SOLVE);
            CU_TRY(launch_chol_solve(sp, smap, h->sm_count, h->stream));
        }
        h->launches++;
        h->work_factored += wn;
        return BP_OK;
    };
    int w_solved = 0;          // windows already solved by the pipelined path (full waves, between the upload segments)
    if (pipelined) {
        if ((rc = run_block_precompute(h, B, 1))) return rc;          // daily blocks do not wait for the bars
        int w_done = 0, k_done[2] = {0, 0};
        for (int s = h->seg_waited; s < h->n_seg; ++s) {
            CU_TRY(cudaStreamWaitEvent(h->stream, h->ev_seg[s], 0));
            const bool last = s == h->n_seg - 1;
            const long long r_end = h->seg_end[s];
            launch_log_returns(h->hf_prices, h->N, h->lr_hf, h->ld, r_end, h->N, h->sm_count, h->stream, h->lr_hf_done);
            h->launches++;
            CU_TRY(cudaGetLastError());
            h->lr_hf_done = r_end;
            int k_ready[2];
            for (int l = 0; l < 2; ++l) {
                k_ready[l] = B.nblocks[0][l];
                if (!last && l == 0 && B.hf_inner) {
                    // day block k is complete once the bars below starts[k+1] have arrived
                    const long long k = std::upper_bound(B.hf_starts.begin() + 1, B.hf_starts.end(), (int)std::min<long long>(r_end, 0x7fffffff)) -
                                        (B.hf_starts.begin() + 1);
                    k_ready[l] = (int)std::min<long long>(B.nblocks[0][l], std::max<long long>(k, k_done[l]));
                } else if (!last && B.hf_blk[l] > 0) {
                    const long long k = floor_div(r_end - B.hf_off, B.hf_blk[l]) - B.hf_bmin[l];
                    k_ready[l] = (int)std::min<long long>(B.nblocks[0][l], std::max<long long>(k, k_done[l]));
                }
            }
            if (B.hf_inner) {
                launch_gather_rows_indexed(h->lr_hf, h->ld, B.hf_starts_dev, h->R, k_done[0], k_ready[0], h->stream);
                h->launches++;
                CU_TRY(cudaGetLastError());
            }
            if ((rc = run_block_precompute(h, B, 0, k_done, k_ready))) return rc;
            k_done[0] = k_ready[0];
            k_done[1] = k_ready[1];
            int w_end = w_done;
            if (last) w_end = B.W;
            else while (w_end < B.W && (long long)b->hf_hi[w_end] <= r_end) ++w_end;
            if (w_end > w_done) {
                const Chunk cw = chunk_at(c, L, w_done);
                PrepParams pp = prep_params(h, b, B, L, cw, w_done, mode);
                {
                    StageTimer tm(h, BP_STAGE_PREP);
                    CU_TRY(launch_window_prep(pp, w_end - w_done, h->stream, h->D));
                    h->launches += pp.use_band ? 2 : 0;
                }
                h->launches++;
                if ((rc = run_gram(h, gram_params(h, B, L, cw, w_done, w_end - w_done, GRAM_S1), B.resampled))) return rc;
                w_done = w_end;
            }
            // Solve the full waves of windows that are ready: the compute stream would otherwise idle until the next
            // segment arrives (the copy takes longer than the statistics / Gram stages), and after the last segment
            // only the remainder (less than one wave) is left instead of every window.  Full waves cost nothing in
            // occupancy whichever of the copy and the GPU is the bottleneck.
            if (!last) {
                const int wave = chol_wave_windows(h->sm_count);
                // Full waves only -- a launch costs the latency of one factorisation however few windows it has -- except
                // before the LAST segment: whatever is ready then is solved while the last segment is on the bus, so that
                // only the last segment's own windows follow the copy (with the default plan that is a full wave anyway;
                // plans with several short tail segments were measured slower, windows.plan_wave_fractions)
                const bool before_last = s == h->n_seg - 2;
                const int wn = before_last ? w_done - w_solved : (w_done - w_solved) / wave * wave;
                if (wn > 0) {
                    if ((rc = solve_range(w_solved, wn))) return rc;
                    w_solved += wn;
                }
            }
        }
        h->seg_waited = h->n_seg;
        h->hf_pending = false;
    }
    for (int w0 = 0; w0 < B.W; w0 += Wc) {
        const int wc = std::min(Wc, B.W - w0);
        const size_t ov = (size_t)w0 * N, om = (size_t)w0 * N * N;
        if (!pipelined) {
            PrepParams pp = prep_params(h, b, B, L, c, w0, mode);
            if (estimator != BP_EST_NONE) pp.beta_den = B.n - 1;       // sample covariance: centre with 1/m, m = n-1 returns
            {
                StageTimer tm(h, BP_STAGE_PREP);
                CU_TRY(launch_window_prep(pp, wc, h->stream, h->D));
                h->launches += pp.use_band ? (mode == BP_MODE_JEFFREYS ? 1 : 2) : 0;
            }
            h->launches++;
        }
        if (out->T) {
            rc = run_gram(h, gram_params(h, B, L, c, w0, wc, GRAM_T), B.resampled);
            if (rc) return rc;
            rc = emit_sym(h, c.S, L, wc, out->T + om);
            if (rc) return rc;
        }
        // pre-summed day blocks: S0 w0, v0, c and rhs are finished from the G w0 partials of the first Gram launch
        // that contracts the intraday phase (it must precede the solve and every emit of rhs / scalars)
        bool post_pending = mode == BP_MODE_CONJUGATE && B.presum;
        auto run_post = [&]() -> int {
            if (!post_pending) return BP_OK;
            post_pending = false;
            PostParams q{};
            q.n_windows = wc;
            q.n_assets = N;
            q.ldv = L.ldv;
            q.mv_part = c.mv;
            q.w0 = c.w0;
            q.gvec = c.gvec;
            q.t = c.t;
            q.s0w0 = c.s0w0;
            q.rhs = c.rhs;
            q.scal = c.scal;
            StageTimer tm(h, BP_STAGE_PREP);
            CU_TRY(launch_conj_post(q, h->stream));
            h->launches++;
            return BP_OK;
        };
        if ((out->S0 || (post_pending && !solve && !out->S1)) && mode == BP_MODE_CONJUGATE) {
            rc = run_gram(h, gram_params(h, B, L, c, w0, wc, GRAM_S0), B.resampled);
            if (rc) return rc;
            if ((rc = run_post())) return rc;
            rc = emit_sym(h, c.S, L, wc, out->S0 ? out->S0 + om : nullptr);
            if (rc) return rc;
        }
        if (solve || out->S1) {
            if (!pipelined) {
                GramParams gp = gram_params(h, B, L, c, w0, wc, mode == BP_MODE_CONJUGATE ? GRAM_S1 : GRAM_J);
                if (G >= 2) {                      // base windows only
                    gp.w_stride = G;
                    gp.n_windows = (wc + G - 1) / G;
                }
                rc = run_gram(h, gp, B.resampled);
                if (rc) return rc;
                if ((rc = run_post())) return rc;
            }
            if (estimator == BP_EST_SHRINKAGE) {
                // C = X_c'X_c  ->  m Sigma_LW / (1 - shrinkage) = C + rho I, rhs = t / (1 - shrinkage), in place (:727-729)
                ShrinkParams q{};
                q.n_windows = wc;
                q.n_assets = N;
                q.n_window = B.n;
                q.ld = h->ld;
                q.ldv = L.ldv;
                q.ldS = L.ldS;
                q.win_stride = L.win_stride;
                q.lr_daily = B.resampled ? h->lr_w : h->lr_d;
                q.rf_row = B.resampled ? h->rf_w : h->rf_row;
                q.day_row = B.day_row + w0;
                q.extra_row = B.extra_row ? B.extra_row + w0 : nullptr;
                q.span_days = B.span + w0;
                q.t = c.t;
                q.S = c.S;
                q.rhs = c.rhs;
                q.scal = c.scal;
                StageTimer tm(h, BP_STAGE_PREP);
                CU_TRY(launch_lw_shrink(q, h->stream));
                h->launches++;
            }
            rc = emit_sym(h, c.S, L, wc, out->S1 ? out->S1 + om : nullptr);
            if (rc) return rc;
        }
        if (solve && G >= 2) {
            // factorise the base windows, then solve the other windows of each group relative to their base
            if ((rc = solve_range(w0, (wc + G - 1) / G, G))) return rc;
            ChainParams cp{};
            cp.n_windows = wc;
            cp.group = G;
            cp.n_assets = N;
            cp.n_window = B.n;
            cp.estimator = estimator;
            cp.ld = h->ld;
            cp.ldv = L.ldv;
            cp.ldS = L.ldS;
            cp.win_stride = L.win_stride;
            cp.inv_gamma = 1.0 / b->risk_aversion;
            cp.lr_daily = h->lr_d;
            cp.day_row = B.day_row + w0;
            cp.S = c.S;
            cp.t = c.t;
            cp.pvec = c.pvec;
            cp.w1 = c.w1;
            cp.nu = c.nu;
            cp.weights = c.weights;
            cp.scal = c.scal;
            cp.status = c.status;
            {
                StageTimer tm(h, BP_STAGE_CHAIN);
                CU_TRY(launch_jeffreys_chain(cp, h->stream));
            }
            h->launches++;
            h->work_chained += wc - (wc + G - 1) / G;
        }
        if (solve) {
            if (G < 2 && (rc = solve_range(w0 + w_solved, wc - w_solved))) return rc;
            if ((rc = emit_vec(h, c.weights, L, wc, out->weights ? out->weights + ov : nullptr))) return rc;
            if ((rc = emit_vec(h, c.nu, L, wc, out->nu ? out->nu + ov : nullptr))) return rc;
            if ((rc = emit_vec(h, c.w1, L, wc, out->w1 ? out->w1 + ov : nullptr))) return rc;
            if ((rc = emit_raw(h, c.status, sizeof(int) * (size_t)wc, out->status ? out->status + w0 : nullptr))) return rc;
        }
        if ((rc = emit_vec(h, c.t, L, wc, out->t ? out->t + ov : nullptr))) return rc;
        if (mode == BP_MODE_CONJUGATE) {
            if ((rc = emit_vec(h, c.w0, L, wc, out->w0 ? out->w0 + ov : nullptr))) return rc;
        }
        if ((rc = emit_vec(h, c.rhs, L, wc, out->rhs ? out->rhs + ov : nullptr))) return rc;
        if ((rc = emit_raw(h, c.scal, sizeof(double) * (size_t)wc * BP_NSCAL,
                           out->scalars ? out->scalars + (size_t)w0 * BP_NSCAL : nullptr))) return rc;
        // the workspace is reused by the next chunk: host-bound copies must have drained
        if (w0 + wc < B.W && h->need_sync) {
            CU_TRY(cudaStreamSynchronize(h->stream));
            h->need_sync = false;
        }
    }
    return finish(h);
}

}  // namespace

extern "C" {

const char* bp_last_error(void) { return g_err.c_str(); }
int bp_version(void) { return 100; }

int bp_init(int device, bp_handle** out) {
    if (!out) return fail(BP_ERR_INVALID, "out is null");
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(BP_ERR_NO_DEVICE, "no CUDA device visible: libbayes_portfolio has no CPU path");
    }
    if (device < 0 || device >= count) return fail(BP_ERR_INVALID, "device %d out of range (%d visible)", device, count);
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(BP_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    CU_TRY(cudaSetDevice(device));
    bp_handle* h = new bp_handle();
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
        delete h;
        return fail(BP_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    }
    h->encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    if (cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_main, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_hf, cudaEventDisableTiming) != cudaSuccess) {
        delete h;
        return fail(BP_ERR_CUDA, "could not create the copy stream / events");
    }
    *out = h;
    return BP_OK;
}

int bp_destroy(bp_handle* h) {
    if (!h) return BP_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    cudaStreamSynchronize(h->copy_stream);
    free_market(h);
    free_pool(h);
    cudaFree(h->pool_cols);
    cudaStreamDestroy(h->copy_stream);
    cudaEventDestroy(h->ev_main);
    cudaEventDestroy(h->ev_hf);
    for (cudaEvent_t e : h->ev_seg)
        if (e) cudaEventDestroy(e);
    for (int k = 0; k < 2; ++k) {
        cudaFree(h->desc_slot[k]);
        cudaFreeHost(h->desc_host_slot[k]);
        if (h->ev_desc[k]) cudaEventDestroy(h->ev_desc[k]);
    }
    for (int a = 0; a < 2; ++a)
        for (int l = 0; l < 2; ++l) cudaFree(h->store[a][l]);
    cudaFree(h->rstore[0]);
    cudaFree(h->rstore[1]);
    cudaFree(h->fstore);
    cudaFree(h->vstore);
    cudaFree(h->prior_n);
    cudaFree(h->band_aw);
    cudaFree(h->band_stats);
    cudaFree(h->ws);
    cudaFree(h->stage);
    for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
    if (h->ev_t0) cudaEventDestroy(h->ev_t0);
    delete h;
    return BP_OK;
}

int bp_set_stream(bp_handle* h, void* cuda_stream) {
    if (!h) return fail(BP_ERR_INVALID, "null handle");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
    if (st != h->stream) {
        // work queued on the old stream still uses the shared workspace, staging buffer and descriptor slots
        CU_TRY(cudaSetDevice(h->device));
        CU_TRY(cudaStreamSynchronize(h->stream));
        h->need_sync = false;
        h->stream = st;
    }
    return BP_OK;
}

int bp_synchronize(bp_handle* h) {
    if (!h) return fail(BP_ERR_INVALID, "null handle");
    CU_TRY(cudaStreamSynchronize(h->copy_stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    h->need_sync = false;
    return BP_OK;
}

int bp_wait_upload(bp_handle* h) {
    if (!h) return fail(BP_ERR_INVALID, "null handle");
    CU_TRY(cudaSetDevice(h->device));
    if (h->has_market && h->R > 0) CU_TRY(cudaEventSynchronize(h->ev_hf));      // the intraday block (copy stream)
    return BP_OK;
}

int bp_set_async_outputs(bp_handle* h, int enable) {
    if (!h) return fail(BP_ERR_INVALID, "null handle");
    if (!enable && h->need_sync) {
        CU_TRY(cudaStreamSynchronize(h->stream));
        h->need_sync = false;
    }
    h->async_outputs = enable != 0;
    return BP_OK;
}

int bp_set_workspace_limit(bp_handle* h, size_t bytes) {
    if (!h) return fail(BP_ERR_INVALID, "null handle");
    h->ws_limit = bytes;
    return BP_OK;
}

int bp_device_info(bp_handle* h, int* sm_count, size_t* free_bytes, size_t* total_bytes) {
    if (!h) return fail(BP_ERR_INVALID, "null handle");
    CU_TRY(cudaSetDevice(h->device));
    size_t f = 0, t = 0;
    CU_TRY(cudaMemGetInfo(&f, &t));
    if (sm_count) *sm_count = h->sm_count;
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return BP_OK;
}

long long bp_launch_count(bp_handle* h) { return h ? h->launches : 0; }

int bp_get_gram_work(bp_handle* h, double* out4) {
    if (!h || !out4) return fail(BP_ERR_INVALID, "null argument");
    out4[0] = h->work_k_rows;       // window rows contracted on the tensor cores (padded to 8)
    out4[1] = h->work_add_blocks;   // precomputed block tiles added per tile pair
    out4[2] = h->work_pre_rows;     // rows contracted by the block precompute launches
    out4[3] = h->work_full_rows;    // rows a from-scratch contraction of every window would touch
    h->work_k_rows = h->work_add_blocks = h->work_pre_rows = h->work_full_rows = 0;
    return BP_OK;
}

int bp_set_jeffreys_chain(bp_handle* h, int group) {
    if (!h) return fail(BP_ERR_INVALID, "null handle");
    if (group < 0 || group > chain_max_group()) return fail(BP_ERR_INVALID, "group must be in [0, %d]", chain_max_group());
    h->chain_group = group;
    return BP_OK;
}

int bp_get_solve_work(bp_handle* h, double* out2) {
    if (!h || !out2) return fail(BP_ERR_INVALID, "null argument");
    out2[0] = h->work_factored;
    out2[1] = h->work_chained;
    h->work_factored = h->work_chained = 0;
    return BP_OK;
}

int bp_set_hf_presum_min_days(bp_handle* h, int min_days) {
    if (!h) return fail(BP_ERR_INVALID, "null handle");
    if (min_days < 0) return fail(BP_ERR_INVALID, "min_days must be >= 0");
    h->presum_min_days = min_days;
    return BP_OK;
}

int bp_set_reuse_min_windows(bp_handle* h, int min_windows) {
    if (!h) return fail(BP_ERR_INVALID, "null handle");
    h->reuse_min_windows = min_windows;
    return BP_OK;
}

int bp_set_upload_pipeline(bp_handle* h, int segments, long long min_bytes) {
    if (!h) return fail(BP_ERR_INVALID, "null handle");
    if (segments < 1 || segments > bp_handle::MAX_SEG || min_bytes < 0)
        return fail(BP_ERR_INVALID, "segments must be in [1,%d] and min_bytes >= 0", bp_handle::MAX_SEG);
    h->pipe_segments = segments;
    h->pipe_min_bytes = (size_t)min_bytes;
    return BP_OK;
}

int bp_set_upload_fractions(bp_handle* h, int n, const double* cum_fractions) {
    if (!h) return fail(BP_ERR_INVALID, "null handle");
    if (n <= 0 || !cum_fractions) {
        h->n_frac = 0;
        return BP_OK;
    }
    if (n > bp_handle::MAX_SEG) return fail(BP_ERR_INVALID, "at most %d segments", bp_handle::MAX_SEG);
    double prev = 0.0;
    for (int i = 0; i < n; ++i) {
        if (!(cum_fractions[i] >= prev) || cum_fractions[i] > 1.0)
            return fail(BP_ERR_INVALID, "cumulative fractions must be non-decreasing in [0, 1]");
        prev = cum_fractions[i];
    }
    for (int i = 0; i < n; ++i) h->frac[i] = cum_fractions[i];
    h->n_frac = n;
    return BP_OK;
}

int bp_solve_wave_windows(bp_handle* h) {
    if (!h) return 0;
    if (cudaSetDevice(h->device) != cudaSuccess) return 0;      // the cluster occupancy query needs the handle's device
    return chol_wave_windows(h->sm_count);
}

int bp_set_stage_timing(bp_handle* h, int enable) {
    if (!h) return fail(BP_ERR_INVALID, "null handle");
    h->timing = enable != 0;
    return BP_OK;
}

int bp_get_stage_times(bp_handle* h, double* ms, long long* launches) {
    if (!h || !ms || !launches) return fail(BP_ERR_INVALID, "null argument");
    CU_TRY(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < BP_NSTAGE; ++i) {
        ms[i] = 0.0;
        launches[i] = 0;
    }
    static const bool timeline = getenv("BP_TIMELINE") != nullptr;
    for (const auto& sp : h->spans) {
        float t = 0.f;
        CU_TRY(cudaEventElapsedTime(&t, sp.a, sp.b));
        ms[sp.stage] += (double)t;
        launches[sp.stage] += 1;
        if (timeline) {
            float t0 = 0.f;
            if (cudaEventElapsedTime(&t0, h->t0_armed ? h->ev_t0 : h->spans.front().a, sp.a) != cudaSuccess) t0 = -1.f;
            fprintf(stderr, "[timeline] stage %d start %8.3f ms dur %7.3f ms\n", sp.stage, t0, t);
        }
    }
    (void)cudaGetLastError();
    h->spans.clear();
    h->ev_used = 0;
    return BP_OK;
}

int bp_prepare_market(bp_handle* h) {
    if (!h) return fail(BP_ERR_INVALID, "null handle");
    if (!h->has_market) return fail(BP_ERR_STATE, "no market uploaded");
    CU_TRY(cudaSetDevice(h->device));
    int rc = wait_hf(h);
    if (rc) return rc;
    {
        StageTimer tm(h, BP_STAGE_LOGRET);
        launch_log_returns(h->prices, h->N, h->lr_d, h->ld, h->D, h->N, h->sm_count, h->stream);
        h->launches++;
        if (h->R > 0) {
            launch_log_returns(h->hf_prices, h->N, h->lr_hf, h->ld, h->R, h->N, h->sm_count, h->stream);
            h->launches++;
        }
    }
    CU_TRY(cudaGetLastError());
    return BP_OK;
}

// Shared body of bp_upload_market / bp_upload_market_async.  Device buffers are reused when the new
// market fits the allocated capacity; the host arrays are copied in their dense layout with plain 1-D
// copies (full PCIe rate from pinned memory) and padded on the fly by the log-return kernel.  The
// intraday block (by far the largest) goes over a second stream together with its log-return kernel,
// so stages that do not read it (Jeffreys, daily statistics) overlap the transfer.
// Working-market buffers for N columns, D daily rows, R intraday rows (grown when needed) and the bookkeeping every
// way of filling them shares (host upload, device gather from the pool)
static int begin_market(bp_handle* h, int N, int D, long long R, int n_mcm) {
    const int ld = round_up(N, 16);
    const size_t need_daily = (size_t)D * ld, need_hf = (size_t)R * ld, need_mcm = (size_t)std::max(n_mcm, 1) * D;
    if (need_daily > h->cap_daily || need_hf > h->cap_hf || need_mcm > h->cap_mcm || (size_t)D > h->cap_days) {
        // grow: everything in flight must be done with the old buffers
        CU_TRY(cudaStreamSynchronize(h->stream));
        CU_TRY(cudaStreamSynchronize(h->copy_stream));
        free_market(h);
        CU_TRY(cudaMalloc(&h->prices, sizeof(double) * need_daily));
        CU_TRY(cudaMalloc(&h->lr_d, sizeof(double) * need_daily));
        CU_TRY(cudaMalloc(&h->caps, sizeof(double) * need_daily));
        CU_TRY(cudaMalloc(&h->rf_row, sizeof(double) * (size_t)D));
        CU_TRY(cudaMalloc(&h->mcm, sizeof(double) * need_mcm));
        if (need_hf) {
            CU_TRY(cudaMalloc(&h->hf_prices, sizeof(double) * need_hf));
            h->lr_hf_ld_cap = ld;
            h->lr_hf_rows_cap = R + R / 8 + 64;        // + room for the gathered overnight rows
            CU_TRY(cudaMalloc(&h->lr_hf, sizeof(double) * (size_t)h->lr_hf_rows_cap * ld));
        }
        h->cap_daily = need_daily;
        h->cap_hf = need_hf;
        h->cap_mcm = need_mcm;
        h->cap_days = (size_t)D;
    }
    if (h->lr_w) {
        // resampled rows refer to the previous prices
        CU_TRY(cudaStreamSynchronize(h->stream));
        free_resampled(h);
    }
    h->N = N; h->D = D; h->ld = ld; h->R = R; h->n_mcm = n_mcm;
    if (h->timing) {
        if (!h->ev_t0) CU_TRY(cudaEventCreate(&h->ev_t0));
        CU_TRY(cudaEventRecord(h->ev_t0, h->stream));
        h->t0_armed = true;
    }
    h->n_seg = 0;
    h->seg_waited = 0;
    h->lr_hf_done = 0;
    h->hf_pending = false;
    return BP_OK;
}

// daily log returns and the tensor maps of the (filled or being filled) working market
static int finish_market(bp_handle* h) {
    const int N = h->N, D = h->D, ld = h->ld;
    const long long R = h->R;
    int rc;
    launch_log_returns(h->prices, N, h->lr_d, ld, D, N, h->sm_count, h->stream);
    h->launches++;
    CU_TRY(cudaGetLastError());
    if ((rc = make_map(h, &h->map_d, h->lr_d, D, ld))) return rc;
    if (R > 0) {
        // rows available behind the R return rows at the CURRENT leading dimension
        h->hf_extra_rows = (long long)((size_t)h->lr_hf_rows_cap * h->lr_hf_ld_cap / ld) - R;
        if (h->hf_extra_rows < 0) h->hf_extra_rows = 0;
        if ((rc = make_map(h, &h->map_hf, h->lr_hf, R + h->hf_extra_rows, ld))) return rc;
    } else {
        h->map_hf = h->map_d;
    }
    h->has_market = true;
    return BP_OK;
}

static int upload_market_impl(bp_handle* h, const bp_market_desc* m, bool blocking) {
    if (!h || !m) return fail(BP_ERR_INVALID, "null handle or market");
    if (m->n_assets <= 0 || m->n_days < 2 || !m->prices || !m->rf_row)
        return fail(BP_ERR_INVALID, "market needs n_assets > 0, n_days >= 2, prices and rf_row");
    if (m->n_hf_rows < 0 || (m->n_hf_rows > 0 && !m->hf_prices)) return fail(BP_ERR_INVALID, "hf_prices missing");
    if (m->n_hf_rows > 0x7fffffffLL) return fail(BP_ERR_INVALID, "too many intraday rows for int32 row indices");
    if (m->n_mcm < 0 || (m->n_mcm > 0 && !m->mcm)) return fail(BP_ERR_INVALID, "mcm missing");
    CU_TRY(cudaSetDevice(h->device));
    const int N = m->n_assets, D = m->n_days;
    const long long R = m->n_hf_rows;
    int rc;
    if ((rc = begin_market(h, N, D, R, m->n_mcm))) return rc;
    h->has_caps = m->caps != nullptr;
    cudaStream_t st = h->stream;
    // the copy stream must not overwrite buffers that earlier launches on the compute stream still read
    CU_TRY(cudaEventRecord(h->ev_main, st));
    CU_TRY(cudaStreamWaitEvent(h->copy_stream, h->ev_main, 0));
    // small daily arrays first: both streams share one host->device copy engine, which serves copies in
    // issue order, and the Jeffreys / daily stages must not queue behind the 1.6 GB intraday block
    CU_TRY(cudaMemcpyAsync(h->prices, m->prices, sizeof(double) * (size_t)D * N, cudaMemcpyHostToDevice, st));
    if (m->caps) CU_TRY(cudaMemcpyAsync(h->caps, m->caps, sizeof(double) * (size_t)D * N, cudaMemcpyHostToDevice, st));
    if (m->n_mcm > 0)
        CU_TRY(cudaMemcpyAsync(h->mcm, m->mcm, sizeof(double) * (size_t)m->n_mcm * D, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(h->rf_row, m->rf_row, sizeof(double) * (size_t)D, cudaMemcpyHostToDevice, st));
    if (R > 0) {
        // asynchronous uploads of a large block go in segments with an event each; the log returns of a segment
        // are computed on the compute stream by its first consumer (wait_hf / the pipelined conjugate path)
        int nseg = 1;
        if (!blocking && sizeof(double) * (size_t)R * N >= h->pipe_min_bytes && R >= 64)
            nseg = h->n_frac > 0 ? h->n_frac : h->pipe_segments;
        // geometric segments (1/2, 1/4, ... of the rows): whatever the compute stream is busy with when the copy
        // starts, little work is left to do after the LAST segment has arrived
        long long r1 = 0;
        for (int s = 0; s < nseg; ++s) {
            const long long r0 = r1;
            r1 = s == nseg - 1 ? R : std::max(r0, R - (R >> (s + 1)));
            if (h->n_frac > 0 && s < nseg - 1)
                r1 = std::min<long long>(R, std::max<long long>(r0, (long long)std::ceil(h->frac[s] * (double)R)));
            CU_TRY(cudaMemcpyAsync(h->hf_prices + (size_t)r0 * N, m->hf_prices + (size_t)r0 * N, sizeof(double) * (size_t)(r1 - r0) * N,
                                   cudaMemcpyHostToDevice, h->copy_stream));
            if (nseg > 1) {
                if (!h->ev_seg[s]) CU_TRY(cudaEventCreateWithFlags(&h->ev_seg[s], cudaEventDisableTiming));
                CU_TRY(cudaEventRecord(h->ev_seg[s], h->copy_stream));
                h->seg_end[s] = r1;
            }
        }
        h->n_seg = nseg > 1 ? nseg : 0;
        CU_TRY(cudaEventRecord(h->ev_hf, h->copy_stream));
        h->hf_pending = true;
    }
    if ((rc = finish_market(h))) return rc;
    if (blocking) {
        // the caller's host arrays may be pageable and may be freed after return
        CU_TRY(cudaStreamSynchronize(h->copy_stream));
        CU_TRY(cudaStreamSynchronize(st));
    }
    return BP_OK;
}

int bp_path_metrics(bp_handle* h, int n_paths, int n_obs, const double* returns, const double* excess, double years,
                    double* out) {
    if (!h || !returns || !excess || !out) return fail(BP_ERR_INVALID, "null handle or array");
    if (n_paths <= 0 || n_obs < 2) return fail(BP_ERR_INVALID, "need at least one series of two returns");
    if (!(years > 0.0)) return fail(BP_ERR_INVALID, "years must be positive");
    CU_TRY(cudaSetDevice(h->device));
    const size_t n = (size_t)n_paths * n_obs;
    int rc = ensure_stage(h, sizeof(double) * (2 * n + (size_t)n_paths * BP_PM_COUNT));
    if (rc) return rc;
    double* d = reinterpret_cast<double*>(h->stage);
    cudaStream_t st = h->stream;
    CU_TRY(cudaMemcpyAsync(d, returns, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(d + n, excess, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    PathMetricsParams p{};
    p.n_paths = n_paths; p.n_obs = n_obs; p.ld = n_obs;
    p.returns = d; p.excess = d + n; p.years = years; p.out = d + 2 * n;
    CU_TRY(launch_path_metrics(p, st));
    h->launches++;
    CU_TRY(cudaMemcpyAsync(out, d + 2 * n, sizeof(double) * (size_t)n_paths * BP_PM_COUNT, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    return BP_OK;
}

int bp_upload_pool(bp_handle* h, const bp_market_desc* m) {
    if (!h) return fail(BP_ERR_INVALID, "null handle");
    CU_TRY(cudaSetDevice(h->device));
    CU_TRY(cudaStreamSynchronize(h->stream));            // earlier gathers may still read the old pool
    free_pool(h);
    if (!m) return BP_OK;                                  // a null market releases the pool
    if (m->n_assets <= 0 || m->n_days < 2 || !m->prices || !m->rf_row)
        return fail(BP_ERR_INVALID, "pool needs n_assets > 0, n_days >= 2, prices and rf_row");
    if (m->n_hf_rows < 0 || (m->n_hf_rows > 0 && !m->hf_prices)) return fail(BP_ERR_INVALID, "hf_prices missing");
    if (m->n_hf_rows > 0x7fffffffLL) return fail(BP_ERR_INVALID, "too many intraday rows for int32 row indices");
    const size_t N = (size_t)m->n_assets, D = (size_t)m->n_days, R = (size_t)m->n_hf_rows;
    cudaStream_t st = h->stream;
    CU_TRY(cudaMalloc(&h->pool_prices, sizeof(double) * D * N));
    CU_TRY(cudaMalloc(&h->pool_rf, sizeof(double) * D));
    CU_TRY(cudaMemcpyAsync(h->pool_prices, m->prices, sizeof(double) * D * N, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(h->pool_rf, m->rf_row, sizeof(double) * D, cudaMemcpyHostToDevice, st));
    if (m->caps) {
        CU_TRY(cudaMalloc(&h->pool_caps, sizeof(double) * D * N));
        CU_TRY(cudaMemcpyAsync(h->pool_caps, m->caps, sizeof(double) * D * N, cudaMemcpyHostToDevice, st));
    }
    if (R > 0) {
        CU_TRY(cudaMalloc(&h->pool_hf, sizeof(double) * R * N));
        CU_TRY(cudaMemcpyAsync(h->pool_hf, m->hf_prices, sizeof(double) * R * N, cudaMemcpyHostToDevice, st));
    }
    CU_TRY(cudaStreamSynchronize(st));                    // the host arrays may be pageable / freed after return
    h->pool_N = m->n_assets;
    h->pool_D = m->n_days;
    h->pool_R = m->n_hf_rows;
    return BP_OK;
}

int bp_select_market(bp_handle* h, const bp_pool_select* s) {
    if (!h || !s) return fail(BP_ERR_INVALID, "null handle or selection");
    if (!h->pool_prices) return fail(BP_ERR_STATE, "no pool resident: call bp_upload_pool first");
    if (s->n_cols <= 0 || !s->cols) return fail(BP_ERR_INVALID, "selection needs at least one column");
    if (s->day_lo < 0 || s->day_hi > h->pool_D || s->day_hi - s->day_lo < 2)
        return fail(BP_ERR_INVALID, "daily row range [%d, %d) outside the pool's %d rows (or shorter than 2)", s->day_lo, s->day_hi, h->pool_D);
    if (s->hf_lo < 0 || s->hf_hi < s->hf_lo || s->hf_hi > h->pool_R)
        return fail(BP_ERR_INVALID, "intraday row range outside the pool");
    for (int j = 0; j < s->n_cols; ++j)
        if (s->cols[j] < 0 || s->cols[j] >= h->pool_N) return fail(BP_ERR_INVALID, "column %d outside the pool's %d columns", s->cols[j], h->pool_N);
    CU_TRY(cudaSetDevice(h->device));
    const int N = s->n_cols, D = s->day_hi - s->day_lo;
    const long long R = s->hf_hi - s->hf_lo;
    int rc;
    if ((rc = begin_market(h, N, D, R, 0))) return rc;
    h->has_caps = h->pool_caps != nullptr;
    cudaStream_t st = h->stream;
    if ((size_t)N > h->pool_cols_cap) {
        CU_TRY(cudaStreamSynchronize(st));
        cudaFree(h->pool_cols);
        h->pool_cols = nullptr;
        h->pool_cols_cap = 0;
        CU_TRY(cudaMalloc(&h->pool_cols, sizeof(int) * (size_t)N));
        h->pool_cols_cap = (size_t)N;
    }
    // (pageable host memory: the copy is staged before the call returns, so the caller's array may go away; stream
    // order keeps earlier gathers, which read the previous column list, ahead of it)
    CU_TRY(cudaMemcpyAsync(h->pool_cols, s->cols, sizeof(int) * (size_t)N, cudaMemcpyHostToDevice, st));
    const long long pN = h->pool_N;
    launch_gather_cols(h->pool_prices + (size_t)s->day_lo * pN, pN, h->prices, N, D, h->pool_cols, h->sm_count, st);
    h->launches++;
    if (h->pool_caps) {
        launch_gather_cols(h->pool_caps + (size_t)s->day_lo * pN, pN, h->caps, N, D, h->pool_cols, h->sm_count, st);
        h->launches++;
    }
    CU_TRY(cudaMemcpyAsync(h->rf_row, h->pool_rf + s->day_lo, sizeof(double) * (size_t)D, cudaMemcpyDeviceToDevice, st));
    if (R > 0) {
        launch_gather_cols(h->pool_hf + (size_t)s->hf_lo * pN, pN, h->hf_prices, N, R, h->pool_cols, h->sm_count, st);
        h->launches++;
        CU_TRY(cudaEventRecord(h->ev_hf, st));
        h->hf_pending = true;            // the first consumer computes the intraday log returns (wait_hf)
    }
    CU_TRY(cudaGetLastError());
    return finish_market(h);
}

int bp_upload_market(bp_handle* h, const bp_market_desc* m) { return upload_market_impl(h, m, true); }

int bp_upload_market_async(bp_handle* h, const bp_market_desc* m) { return upload_market_impl(h, m, false); }

int bp_set_resampled(bp_handle* h, const bp_resampled_desc* r) {
    if (!h || !r) return fail(BP_ERR_INVALID, "null argument");
    if (!h->has_market) return fail(BP_ERR_STATE, "no market uploaded");
    if (r->n_rows < 2 || !r->num_row || !r->den_row || !r->rf_row) return fail(BP_ERR_INVALID, "bp_resampled_desc incomplete");
    for (int i = 0; i < r->n_rows; ++i)
        if (r->num_row[i] < 0 || r->num_row[i] >= h->D || r->den_row[i] < 0 || r->den_row[i] >= h->D)
            return fail(BP_ERR_INVALID, "resampled row %d references a price row outside [0,%d)", i, h->D);
    CU_TRY(cudaSetDevice(h->device));
    CU_TRY(cudaStreamSynchronize(h->stream));
    free_resampled(h);
    const int Rw = r->n_rows;
    CU_TRY(cudaMalloc(&h->lr_w, sizeof(double) * (size_t)Rw * h->ld));
    CU_TRY(cudaMalloc(&h->rf_w, sizeof(double) * (size_t)Rw));
    CU_TRY(cudaMalloc(&h->rs_idx, sizeof(int) * 2 * (size_t)Rw));
    CU_TRY(cudaMemcpyAsync(h->rs_idx, r->num_row, sizeof(int) * (size_t)Rw, cudaMemcpyHostToDevice, h->stream));
    CU_TRY(cudaMemcpyAsync(h->rs_idx + Rw, r->den_row, sizeof(int) * (size_t)Rw, cudaMemcpyHostToDevice, h->stream));
    CU_TRY(cudaMemcpyAsync(h->rf_w, r->rf_row, sizeof(double) * (size_t)Rw, cudaMemcpyHostToDevice, h->stream));
    if (r->mcm && h->n_mcm > 0) {
        CU_TRY(cudaMalloc(&h->mcm_w, sizeof(double) * (size_t)h->n_mcm * Rw));
        CU_TRY(cudaMemcpyAsync(h->mcm_w, r->mcm, sizeof(double) * (size_t)h->n_mcm * Rw, cudaMemcpyHostToDevice, h->stream));
    }
    launch_gather_log_returns(h->prices, h->N, h->rs_idx, h->rs_idx + Rw, h->lr_w, h->ld, Rw, h->N, h->stream);
    h->launches++;
    CU_TRY(cudaGetLastError());
    h->Rw = Rw;
    int rc = make_map(h, &h->map_w, h->lr_w, Rw, h->ld);
    if (rc) return rc;
    CU_TRY(cudaStreamSynchronize(h->stream));     // host index arrays may be freed after return
    return BP_OK;
}

int bp_conjugate_batched(bp_handle* h, const bp_window_batch* b, const bp_outputs* out) {
    if (!out) return fail(BP_ERR_INVALID, "outputs missing");
    return run_batches(h, b, out, BP_MODE_CONJUGATE, true);
}

int bp_jeffreys_batched(bp_handle* h, const bp_window_batch* b, const bp_outputs* out) {
    if (!out) return fail(BP_ERR_INVALID, "outputs missing");
    return run_batches(h, b, out, BP_MODE_JEFFREYS, true);
}

int bp_estimator_batched(bp_handle* h, const bp_window_batch* b, int estimator, const bp_outputs* out) {
    if (!h || !b || !out) return fail(BP_ERR_INVALID, "null argument");
    if (estimator != BP_EST_JORION && estimator != BP_EST_SHRINKAGE)
        return fail(BP_ERR_INVALID, "estimator must be BP_ESTIMATOR_JORION or BP_ESTIMATOR_SHRINKAGE");
    if (out->T || out->S0) return fail(BP_ERR_INVALID, "estimator batches do not return T / S0 (use bp_stats_batched)");
    return run_batches(h, b, out, BP_MODE_JEFFREYS, true, estimator);
}

int bp_stats_batched(bp_handle* h, const bp_window_batch* b, double* t, double* T) {
    bp_outputs o;
    memset(&o, 0, sizeof(o));
    o.t = t;
    o.T = T;
    return run_batches(h, b, &o, BP_MODE_JEFFREYS, false);
}

int bp_hf_cov_batched(bp_handle* h, const bp_window_batch* b, double* n0, double* S0) {
    if (!h || !b) return fail(BP_ERR_INVALID, "null handle or batch");
    bp_outputs o;
    memset(&o, 0, sizeof(o));
    o.S0 = S0;
    std::vector<double> scal;
    if (n0) {
        scal.resize((size_t)std::max(b->n_windows, 0) * BP_NSCAL);
        o.scalars = scal.data();
    }
    int rc = run_batches(h, b, &o, BP_MODE_CONJUGATE, false);
    if (rc) return rc;
    if (n0) {
        if (is_device_ptr(n0)) return fail(BP_ERR_INVALID, "bp_hf_cov_batched: n0 must be a host pointer");
        for (int w = 0; w < b->n_windows; ++w) n0[w] = scal[(size_t)w * BP_NSCAL + BP_SCAL_N0];
    }
    return BP_OK;
}

int bp_moments_batched(bp_handle* h, const bp_window_batch* b, int mode, const bp_outputs* out) {
    if (!out) return fail(BP_ERR_INVALID, "outputs missing");
    if (mode != BP_MODE_CONJUGATE && mode != BP_MODE_JEFFREYS) return fail(BP_ERR_INVALID, "mode must be 0 or 1");
    if (out->weights || out->nu || out->w1 || out->status)
        return fail(BP_ERR_INVALID, "bp_moments_batched does not solve: use bp_conjugate_batched / bp_jeffreys_batched");
    return run_batches(h, b, out, mode, false);
}

int bp_backtest_batched(bp_handle* h, const bp_backtest_desc* d) {
    if (!h || !d) return fail(BP_ERR_INVALID, "null argument");
    if (!h->has_market) return fail(BP_ERR_STATE, "no market uploaded");
    if (!h->has_caps) return fail(BP_ERR_STATE, "the backtest loop needs market caps (comparison portfolio)");
    const int R = d->n_rebalances, N = h->N;
    if (R < 1 || !d->reb_row || !d->weights || !d->returns || !d->turnover || !d->metrics)
        return fail(BP_ERR_INVALID, "bp_backtest_desc incomplete");
    if (N > 8 * 256) return fail(BP_ERR_INVALID, "the loop kernel supports at most 2048 assets");
    if (is_device_ptr(d->reb_row)) return fail(BP_ERR_INVALID, "reb_row must be a host pointer");
    for (int s = 0; s < R; ++s) {
        const int r = d->reb_row[s];
        if (r < 1 || r >= h->D || (s > 0 && r <= d->reb_row[s - 1]))
            return fail(BP_ERR_INVALID, "reb_row must be ascending rows in [1, n_days)");
    }
    CU_TRY(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    if (d->last_row < d->reb_row[R - 1] || d->last_row >= h->D) return fail(BP_ERR_INVALID, "last_row must lie in [reb_row[R-1], n_days)");
    const int T = d->last_row - d->reb_row[0];
    // workspace: weights, member, reb_row, returns, turnover, metrics
    auto al = [](size_t x) { return (x + 15) & ~(size_t)15; };
    const size_t b_w = al(sizeof(double) * (size_t)R * N), b_m = al((size_t)R * N), b_r = al(sizeof(int) * (size_t)R);
    const size_t b_ret = al(sizeof(double) * (size_t)std::max(T, 1)), b_to = al(sizeof(double) * (size_t)std::max(R - 1, 1));
    const size_t b_met = al(sizeof(double) * (size_t)R * 5);
    int rc = ensure_ws(h, b_w + b_m + b_r + b_ret + b_to + b_met + 256);
    if (rc) return rc;
    unsigned char* base = h->ws;
    double* dw = reinterpret_cast<double*>(base);                    base += b_w;
    unsigned char* dm = base;                                          base += b_m;
    int* dr = reinterpret_cast<int*>(base);                           base += b_r;
    double* dret = reinterpret_cast<double*>(base);                   base += b_ret;
    double* dto = reinterpret_cast<double*>(base);                    base += b_to;
    double* dmet = reinterpret_cast<double*>(base);
    const double* wsrc = d->weights;
    if (!is_device_ptr(d->weights)) {
        CU_TRY(cudaMemcpyAsync(dw, d->weights, sizeof(double) * (size_t)R * N, cudaMemcpyHostToDevice, st));
        wsrc = dw;
    }
    const unsigned char* msrc = d->member;
    if (d->member && !is_device_ptr(d->member)) {
        CU_TRY(cudaMemcpyAsync(dm, d->member, (size_t)R * N, cudaMemcpyHostToDevice, st));
        msrc = dm;
    }
    CU_TRY(cudaMemcpyAsync(dr, d->reb_row, sizeof(int) * (size_t)R, cudaMemcpyHostToDevice, st));
    LoopParams lp{};
    lp.n_assets = N; lp.n_rebalances = R; lp.ldw = N;
    lp.weights = wsrc; lp.member = msrc; lp.reb_row = dr; lp.last_row = d->last_row;
    lp.prices = h->prices; lp.ld_prices = N; lp.caps = h->caps; lp.ld_caps = N; lp.rf_row = h->rf_row;
    lp.distance_scale = d->distance_scale; lp.turnover_cost_bps = d->turnover_cost_bps;
    lp.returns = dret; lp.turnover = dto; lp.metrics = dmet;
    CU_TRY(launch_backtest_loop(lp, st));
    h->launches++;
    if (T > 0 && (rc = emit_raw(h, dret, sizeof(double) * (size_t)T, d->returns))) return rc;
    if (R > 1 && (rc = emit_raw(h, dto, sizeof(double) * (size_t)(R - 1), d->turnover))) return rc;
    if ((rc = emit_raw(h, dmet, sizeof(double) * (size_t)R * 5, d->metrics))) return rc;
    CU_TRY(cudaStreamSynchronize(st));       // the staging copies of host inputs must be consumed before return
    h->need_sync = false;
    return BP_OK;
}

int bp_excess_returns(bp_handle* h, const bp_window_batch* b, double* X) {
    if (!h || !b || !X) return fail(BP_ERR_INVALID, "null argument");
    if (b->n_windows != 1) return fail(BP_ERR_INVALID, "bp_excess_returns handles one window per call");
    Batch B;
    CU_TRY(cudaSetDevice(h->device));
    int rc = upload_batch(h, b, false, &B);
    if (rc) return rc;
    const int K = b->rolling_window - 1;
    const size_t bytes = sizeof(double) * (size_t)K * h->N;
    const bool dev = is_device_ptr(X);
    double* dst = X;
    if (!dev) {
        if ((rc = ensure_stage(h, bytes))) return rc;
        dst = reinterpret_cast<double*>(h->stage);
    }
    launch_excess_returns(h->lr_d, h->ld, h->rf_row, b->day_row[0], b->span_days[0], b->rolling_window, h->N, dst, h->stream);
    h->launches++;
    CU_TRY(cudaGetLastError());
    if (!dev) {
        CU_TRY(cudaMemcpyAsync(X, dst, bytes, cudaMemcpyDeviceToHost, h->stream));
        CU_TRY(cudaStreamSynchronize(h->stream));
    }
    return BP_OK;
}

int bp_quadratic_form(bp_handle* h, int n, const double* w, const double* S, double* out) {
    if (!h || !w || !S || !out || n <= 0) return fail(BP_ERR_INVALID, "bad argument");
    CU_TRY(cudaSetDevice(h->device));
    const size_t need = sizeof(double) * ((size_t)n * n + n + 8);
    int rc = ensure_stage(h, need);
    if (rc) return rc;
    double* dS = reinterpret_cast<double*>(h->stage);
    double* dw = dS + (size_t)n * n;
    double* dv = dw + n;
    CU_TRY(cudaMemcpyAsync(dS, S, sizeof(double) * (size_t)n * n, cudaMemcpyHostToDevice, h->stream));
    CU_TRY(cudaMemcpyAsync(dw, w, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    launch_quadform(dS, n, dw, n, dv, 0.0, 0.0, nullptr, nullptr, h->stream);
    h->launches++;
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(out, dv, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    return BP_OK;
}

int bp_dense_posterior(bp_handle* h, const bp_dense_problem* in, const bp_dense_result* out) {
    if (!h || !in || !out) return fail(BP_ERR_INVALID, "null argument");
    const int N = in->n_assets;
    const bool c_only = !in->T && !in->t && !in->jeffreys;
    if (N <= 0 || (!c_only && (!in->T || !in->t))) return fail(BP_ERR_INVALID, "T and t are required");
    if (!in->jeffreys && (!in->S0 || !in->w0)) return fail(BP_ERR_INVALID, "conjugate posterior needs S0 and w0");
    if (!(in->risk_aversion != 0.0)) return fail(BP_ERR_INVALID, "risk_aversion must be non-zero");
    CU_TRY(cudaSetDevice(h->device));
    const int ldS = round_up(N, 32), rowsS = ldS + 8, ldv = round_up(N, 16);
    const size_t NN = (size_t)N * N;
    auto al = [](size_t x) { return (x + 3) & ~(size_t)3; };      // keep every carve 32-byte aligned
    const size_t doubles = 3 * al(NN) + 4 * al(N) + 4 * al(ldv) + (size_t)rowsS * ldS + al(BP_NSCAL) + 8;
    int rc = ensure_ws(h, sizeof(double) * doubles + 64);
    if (rc) return rc;
    double* p = reinterpret_cast<double*>(h->ws);
    double* dT = p;            p += al(NN);
    double* dS0 = p;           p += al(NN);
    double* dS1in = p;         p += al(NN);
    double* dt = p;            p += al(N);
    double* dw0 = p;           p += al(N);
    double* dw1in = p;         p += al(N);
    double* ds0w0 = p;         p += al(N);
    double* drhs = p;          p += al(ldv);
    double* dw1 = p;           p += al(ldv);
    double* dnu = p;           p += al(ldv);
    double* dwts = p;          p += al(ldv);
    double* dscal = p;         p += al(BP_NSCAL);
    double* dS = p;            p += (size_t)rowsS * ldS;
    int* dstatus = reinterpret_cast<int*>(p);
    cudaStream_t st = h->stream;
    CU_TRY(cudaMemsetAsync(dscal, 0, sizeof(double) * BP_NSCAL, st));
    CU_TRY(cudaMemsetAsync(dstatus, 0, sizeof(int), st));
    if (!c_only) {
        CU_TRY(cudaMemcpyAsync(dT, in->T, sizeof(double) * NN, cudaMemcpyHostToDevice, st));
        CU_TRY(cudaMemcpyAsync(dt, in->t, sizeof(double) * N, cudaMemcpyHostToDevice, st));
    }
    if (!in->jeffreys) {
        CU_TRY(cudaMemcpyAsync(dS0, in->S0, sizeof(double) * NN, cudaMemcpyHostToDevice, st));
        CU_TRY(cudaMemcpyAsync(dw0, in->w0, sizeof(double) * N, cudaMemcpyHostToDevice, st));
        if (in->S1) CU_TRY(cudaMemcpyAsync(dS1in, in->S1, sizeof(double) * NN, cudaMemcpyHostToDevice, st));
        if (in->w1) CU_TRY(cudaMemcpyAsync(dw1in, in->w1, sizeof(double) * N, cudaMemcpyHostToDevice, st));
    }
    const double n1 = in->n1 ? *in->n1 : in->n0 + (double)in->rolling_window;
    DenseParams dp{};
    dp.n_assets = N; dp.ldS = ldS; dp.ldv = ldv; dp.n_window = in->rolling_window;
    dp.n0 = in->n0; dp.n1 = n1; dp.has_c = in->c != nullptr; dp.c_in = in->c ? *in->c : 0.0;
    dp.T = c_only ? nullptr : dT; dp.t = c_only ? nullptr : dt; dp.S0 = dS0; dp.w0 = dw0; dp.S1_in = (!in->jeffreys && in->S1) ? dS1in : nullptr;
    dp.s0w0 = ds0w0; dp.rhs = drhs; dp.S_out = dS; dp.scal = dscal;
    launch_dense_prep(dp, in->jeffreys != 0, st);
    h->launches++;
    CU_TRY(cudaGetLastError());
    if (c_only) {
        if (out->scalars) CU_TRY(cudaMemcpyAsync(out->scalars, dscal, sizeof(double) * BP_NSCAL, cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        return BP_OK;
    }
    if (out->S1) {
        const Layout L{N, ldv, ldS, rowsS, (long long)rowsS * ldS, 2, 0};
        if ((rc = emit_sym(h, dS, L, 1, out->S1))) return rc;
        if (h->need_sync) { CU_TRY(cudaStreamSynchronize(st)); h->need_sync = false; }
    }
    const double inv_gamma = 1.0 / in->risk_aversion;
    if (!in->jeffreys && in->w1) {
        // injected posterior w: v1 = w1'S1w1 (:574), nu (:572-575), weights (:836)
        launch_quadform(dS, ldS, dw1in, N, dscal + BP_SCAL_V1, n1, inv_gamma, dnu, dwts, st);
        h->launches++;
        CU_TRY(cudaMemcpyAsync(dw1, dw1in, sizeof(double) * N, cudaMemcpyDeviceToDevice, st));
    } else {
        SolveParams sp{};
        sp.n_windows = 1; sp.n_assets = N; sp.ldS = ldS; sp.win_stride = (long long)rowsS * ldS; sp.ldv = ldv;
        sp.mode = in->jeffreys ? BP_MODE_JEFFREYS : BP_MODE_CONJUGATE;
        sp.inv_gamma = inv_gamma;
        sp.S = dS; sp.rhs = drhs; sp.scal = dscal; sp.w1 = dw1; sp.nu = dnu; sp.weights = dwts; sp.status = dstatus;
        CUtensorMap smap;
        if ((rc = make_solve_map(h, &smap, dS, rowsS, ldS))) return rc;
        CU_TRY(launch_chol_solve(sp, smap, h->sm_count, st));
        h->launches++;
    }
    CU_TRY(cudaGetLastError());
    if (out->scalars) CU_TRY(cudaMemcpyAsync(out->scalars, dscal, sizeof(double) * BP_NSCAL, cudaMemcpyDeviceToHost, st));
    if (out->w1) CU_TRY(cudaMemcpyAsync(out->w1, dw1, sizeof(double) * N, cudaMemcpyDeviceToHost, st));
    if (out->nu) CU_TRY(cudaMemcpyAsync(out->nu, dnu, sizeof(double) * N, cudaMemcpyDeviceToHost, st));
    if (out->weights) CU_TRY(cudaMemcpyAsync(out->weights, dwts, sizeof(double) * N, cudaMemcpyDeviceToHost, st));
    if (out->status) CU_TRY(cudaMemcpyAsync(out->status, dstatus, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    return BP_OK;
}

}  // extern "C"
