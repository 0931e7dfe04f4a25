// Backtest loop body on the device (portfolio_calculations.py:1054-1075, :1077-1104, :1127-1219):
// daily portfolio return, weight drift, turnover, transaction cost and the weight metrics, for all
// rebalance segments of a backtest at once.  One CTA per segment (the days after rebalance s up to and
// including rebalance s+1): weights only couple consecutive days inside a segment (:1148-1159), and the
// new weights at a rebalance date never depend on the old ones, so segments are independent; with daily
// rebalancing every trading day is its own segment.
#include "common.cuh"
#include "kernels.h"

namespace bp {

constexpr int LOOP_THREADS = 256;

__global__ void __launch_bounds__(LOOP_THREADS) backtest_loop_kernel(LoopParams p) {
    __shared__ double scratch[40];
    const int s = blockIdx.x;                 // segment index = rebalance index; block R = days after the last rebalance
    const int tid = threadIdx.x;
    const int N = p.n_assets;
    const bool tail = s == p.n_rebalances;
    const double* w_new_ptr = p.weights + (long long)(tail ? s - 1 : s) * p.ldw;
    const int d_reb = tail ? p.last_row : p.reb_row[s];

    // ---- weight metrics of rebalance s (:1194-1209) and distance to the value-weighted portfolio (:1077-1104)
    if (!tail) {
        double mx = -1.0, mn = 1.0, sl = 0.0, ss = 0.0, cl = 0.0, cs = 0.0, capsum = 0.0;
        for (int i = tid; i < N; i += LOOP_THREADS) {
            const double w = w_new_ptr[i];
            if (p.member == nullptr || p.member[(long long)s * N + i]) {
                if (w > 0.0) { mx = fmax(mx, w); sl += w; cl += 1.0; }
                if (w < 0.0) { mn = fmin(mn, w); ss += w; cs += 1.0; }
                capsum += p.caps[(long long)d_reb * p.ld_caps + i];
            }
        }
        // block max / min through sums of one-hot comparisons would lose exactness: use shuffles
        for (int o = 16; o > 0; o >>= 1) {
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        }
        __shared__ double smx[LOOP_THREADS / 32], smn[LOOP_THREADS / 32];
        if ((tid & 31) == 0) { smx[tid >> 5] = mx; smn[tid >> 5] = mn; }
        sl = block_sum(sl, scratch);
        ss = block_sum(ss, scratch);
        cl = block_sum(cl, scratch);
        cs = block_sum(cs, scratch);
        capsum = block_sum(capsum, scratch);
        for (int k = 0; k < LOOP_THREADS / 32; ++k) { mx = fmax(mx, smx[k]); mn = fmin(mn, smn[k]); }
        double dist = 0.0, members = 0.0;
        for (int i = tid; i < N; i += LOOP_THREADS) {
            if (p.member == nullptr || p.member[(long long)s * N + i]) {
                const double vw = p.caps[(long long)d_reb * p.ld_caps + i] / capsum;
                dist += fabs(w_new_ptr[i] * p.distance_scale - vw);
                members += 1.0;
            }
        }
        dist = block_sum(dist, scratch);
        members = block_sum(members, scratch);
        if (tid == 0) {
            const double nan = __longlong_as_double(0x7ff8000000000000LL);
            double* m = p.metrics + (long long)s * 5;
            m[0] = cl > 0.0 ? mx : nan;             // max_long   (max of an empty selection is NaN in pandas)
            m[1] = cs > 0.0 ? mn : nan;             // max_short
            m[2] = cl > 0.0 ? sl / cl : nan;        // avg_long
            m[3] = cs > 0.0 ? ss / cs : nan;        // avg_short
            m[4] = dist / members;                  // average_distance_to_comparison_portfolio
        }
    }
    if (s == 0) return;                       // the first rebalance has no preceding segment

    // ---- segment: days reb_row[s-1]+1 .. reb_row[s], starting from the weights chosen at rebalance s-1.
    // Only the stocks HELD since rebalance s-1 take part (the reference reindexes the day's returns to the
    // portfolio's own index, :1134); columns outside that universe may hold NaN prices (before listing, after
    // delisting) and are never read.  A held stock whose return is NaN drops out of every sum exactly as pandas'
    // skipna sums drop it (:1137, :1143, :1156): its weight becomes NaN and stays NaN for the rest of the segment,
    // and the turnover's outer merge + fillna(0) (:1057-1064) counts it as 0.
    const double* w_old_ptr = p.weights + (long long)(s - 1) * p.ldw;
    const unsigned char* held_ptr = p.member ? p.member + (long long)(s - 1) * N : nullptr;
    constexpr int PER = 8;                    // assets per thread (N <= 2048)
    double w[PER];
    bool held[PER];
#pragma unroll
    for (int e = 0; e < PER; ++e) {
        const int i = tid + e * LOOP_THREADS;
        held[e] = i < N && (held_ptr == nullptr || held_ptr[i] != 0);
        w[e] = held[e] ? w_old_ptr[i] : 0.0;
    }
    for (int d = p.reb_row[s - 1] + 1; d <= d_reb; ++d) {
        double pr = 0.0, sw = 0.0;
        double ret[PER];
#pragma unroll
        for (int e = 0; e < PER; ++e) {
            const int i = tid + e * LOOP_THREADS;
            ret[e] = 0.0;
            if (held[e]) {
                ret[e] = p.prices[(long long)d * p.ld_prices + i] / p.prices[(long long)(d - 1) * p.ld_prices + i] - 1.0;  // pct_change
                const double prod = ret[e] * w[e];
                if (!isnan(prod)) pr += prod;                          // Series.sum() skips NaN (:1137)
                if (!isnan(w[e])) sw += w[e];
            }
        }
        pr = block_sum(pr, scratch);
        sw = block_sum(sw, scratch);
        const double rf_d = pow(p.rf_row[d] + 1.0, 1.0 / 252.0) - 1.0;  // :1142
        double r = pr + (1.0 - sw) * rf_d;                             // :1143
        const double upd_rf = (1.0 - sw) * (1.0 + rf_d);               // :1148-1149
        double sw2 = 0.0;
#pragma unroll
        for (int e = 0; e < PER; ++e) {
            if (held[e]) {
                w[e] = w[e] * (1.0 + ret[e]);                          // :1152-1153
                if (!isnan(w[e])) sw2 += w[e];
            }
        }
        sw2 = block_sum(sw2, scratch);
        const double total = sw2 + upd_rf;                             // :1156
#pragma unroll
        for (int e = 0; e < PER; ++e) w[e] = w[e] / total;             // :1159
        if (d == d_reb && !tail) {
            // turnover against the new weights (:1054-1075) and transaction cost (:1214-1215): outer merge of
            // the two index sets, missing / NaN entries count as 0
            const unsigned char* new_ptr = p.member ? p.member + (long long)s * N : nullptr;
            double diff = 0.0, sb = 0.0, sa = 0.0;
#pragma unroll
            for (int e = 0; e < PER; ++e) {
                const int i = tid + e * LOOP_THREADS;
                if (i < N) {
                    const bool in_new = new_ptr == nullptr || new_ptr[i] != 0;
                    const double wa_raw = in_new ? w_new_ptr[i] : 0.0;
                    const double wa = isnan(wa_raw) ? 0.0 : wa_raw;
                    const double wb = (held[e] && !isnan(w[e])) ? w[e] : 0.0;
                    diff += fabs(wb - wa);
                    sb += wb;
                    sa += wa;
                }
            }
            diff = block_sum(diff, scratch);
            sb = block_sum(sb, scratch);
            sa = block_sum(sa, scratch);
            const double turnover = (diff + fabs(sb - sa)) / 2.0;
            r -= p.turnover_cost_bps / 10000.0 * turnover;
            if (tid == 0) p.turnover[s - 1] = turnover;
        }
        if (tid == 0) p.returns[d - p.reb_row[0] - 1] = r;
    }
}

cudaError_t launch_backtest_loop(const LoopParams& p, cudaStream_t st) {
    if (p.n_rebalances <= 0) return cudaSuccess;
    backtest_loop_kernel<<<p.n_rebalances + 1, LOOP_THREADS, 0, st>>>(p);
    return cudaGetLastError();
}

}  // namespace bp
