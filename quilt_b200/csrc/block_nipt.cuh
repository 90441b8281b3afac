// block_nipt.cuh — block Gibbs for three-haplotype (NIPT, ff > 0) calls.
//
// Reference path (QUILT/src/gibbs-nipt-block.cpp, block_approach = 6, consider_total_relabelling = false,
// resample_H_using_H_class = true — the defaults that production never overrides, gibbs-nipt.cpp:3025):
//   Rcpp_define_blocked_snps_using_gamma_on_the_fly  :311-523   -> k_block_rate + k_block_define
//     (smoothing copied-from-stitch.cpp:446-518, stopping rule :522-567, quantile gibbs-nipt-block.cpp:81-85)
//   Rcpp_make_gibbs_considers                        :1307-1553 -> k_block_define
//   Rcpp_block_gibbs_resampler                       :1636-1967 -> k_block_nipt
//     (Rcpp_gibbs_block_forward_one :1122-1253, Rcpp_consider_block_relabelling :590-949,
//      Rcpp_reset_local_variables :1257-1292, H_class prior :169-208, :251-279)
//   rcpp_sample_H_using_H_class                      :213-246   -> k_sample_H
//   trailing re-forward / backward                   :1900-1958 -> k_make_eG, k_fb_generic (forward), k_bwd_fast
//
// For diploid calls the block resampler is the identity and none of this runs.
#pragma once

#include "device_common.cuh"
#include "prep.cuh"
#include "sweep.cuh"
#include "types.h"

namespace qb {

// rr (gibbs-nipt-block.cpp:1760-1768): permutation ir maps emission label i to haplotype slot RR[ir][i] - 1
__device__ __constant__ int c_RR[6][3] = {{1, 2, 3}, {1, 3, 2}, {2, 1, 3}, {2, 3, 1}, {3, 1, 2}, {3, 2, 1}};
// rx (gibbs-nipt-block.cpp:766-773)
__device__ __constant__ int c_RX[6][3] = {{1, 2, 3}, {1, 3, 2}, {2, 1, 3}, {3, 1, 2}, {2, 3, 1}, {3, 2, 1}};

// rate2[g] = sum_h (1 - sigma_g * sum_k alpha_h[k, g] beta_h[k, g + 1] eMatGrid_h[k, g + 1]), g < T - 2; 0 beyond.
// grid = (T, jobs), 256 threads
__global__ void __launch_bounds__(256) k_block_rate(BatchParams P, const JobDev* __restrict__ jobs) {
    __shared__ double red[3][8];
    const JobDev& J = jobs[blockIdx.y];
    if (*J.underflow) return;
    const int g = blockIdx.x, K = P.K, Kp = P.Kp, T = P.T;
    BlockScratch B;
    B.carve(J.blk, T);
    if (g >= T - 2) {
        if (threadIdx.x == 0 && g < T) B.rate2[g] = 0.0;
        return;
    }
    const int nh = (P.ff > 0) ? 3 : 2;
    double s[3] = {0, 0, 0};
    for (int h = 0; h < nh; h++) {
        const double* a = J.alpha + ((size_t)h * T + g) * Kp;
        const double* b = J.beta + ((size_t)h * T + g + 1) * Kp;
        const double* e = J.eG + ((size_t)h * T + g + 1) * Kp;
        for (int k = threadIdx.x; k < K; k += 256) s[h] += (ld_stream(a + k) * ld_stream(b + k)) * ld_stream(e + k);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int h = 0; h < 3; h++) {
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) s[h] += __shfl_xor_sync(0xffffffffu, s[h], d);
        if (lane == 0) red[h][warp] = s[h];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double d = J.tm[2 * g];
        double r = 0;
        for (int h = 0; h < nh; h++) {
            double t = 0;
            for (int w = 0; w < 8; w++) t += red[h][w];
            r += 1 - d * t;
        }
        B.rate2[g] = r;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Block definition + "considers" (Rcpp_define_blocked_snps_using_gamma_on_the_fly gibbs-nipt-block.cpp:311-523 with
// rcpp_make_smoothed_rate / rcpp_determine_where_to_stop copied-from-stitch.cpp:446-567, Rcpp_make_gibbs_considers
// gibbs-nipt-block.cpp:1307-1553), as one CTA per job:
//   1. smoothed switch rate: one thread per grid boundary (every window is independent; the additions keep the reference's
//      order: nearest interval first, left side before right side);
//   2. both stable orders of the smoothed rate (ascending for the quantile, descending for the peak list) from ONE
//      all-pairs rank count — no sort;
//   3. peak picking: the only truly serial part (a peak claims its valley and removes the candidates inside it) is walked by
//      one warp: 32 list entries are screened per ballot, the two valley walks evaluate 32 steps at a time (prefix minimum by
//      shuffle scan, first stopping step by ballot), ranges are cleared lane-parallel;
//   4. block ids, block / read ranges, removal of read-less blocks: flags + block-wide prefix counts.
// The outputs (n_blocks, grid_start/_end, reads_start/_end, grid_where) are integers defined by comparisons of doubles that
// are computed with the reference's operation order, so they are identical to the reference's.
constexpr int BD_NT = 256;

// exclusive prefix count of flag[0..n) into out[0..n) (may alias), returns the total; contiguous chunk per thread
__device__ inline int bd_scan_excl(const int32_t* flag, int32_t* out, int n, int* s_part /*[BD_NT / 32]*/) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int chunk = (n + BD_NT - 1) / BD_NT;
    const int lo = min(n, tid * chunk), hi = min(n, lo + chunk);
    int mine = 0;
    for (int i = lo; i < hi; i++) mine += flag[i];
    int incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += y;
    }
    __syncthreads();  // (s_part may still be read from a previous call)
    if (lane == 31) s_part[warp] = incl;
    __syncthreads();
    int base = incl - mine, total = 0;
#pragma unroll
    for (int w = 0; w < BD_NT / 32; w++) {
        const int v = s_part[w];
        if (w < warp) base += v;
        total += v;
    }
    for (int i = lo; i < hi; i++) {
        const int f = flag[i];
        out[i] = base;
        base += f;
    }
    __syncthreads();
    return total;
}

// One side of a peak's valley (copied-from-stitch.cpp:522-567): step c looks at position peak + dir * c; the walk ends at
// the first step that is near an end of the region, whose outer neighbour is no longer available, that has climbed to more
// than three times the smallest rate seen, or that is below the threshold but rising over five steps; the result is the
// position of the smallest rate seen up to and including that step (first occurrence).  Warp-cooperative: lane l evaluates
// step c0 + l + 1; all lanes return the same value.
__device__ inline int bd_valley(const double* rate, const uint8_t* avail, int peak, double thresh, int T, int dir, int lane) {
    double carry_min = rate[peak];
    int carry_arg = peak;
    const double at_peak = carry_min;
    for (int c0 = 0;; c0 += 32) {
        const int c = c0 + lane + 1;
        const int pos = peak + dir * c;
        const bool valid = pos >= 0 && pos <= T - 2;
        const double v = valid ? rate[pos] : __longlong_as_double(0x7ff0000000000000ll);
        // smallest rate over steps c0 + 1 .. c (strict comparisons: the earliest position wins ties), then the carry
        double m = v;
        int arg = pos;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double om = __shfl_up_sync(0xffffffffu, m, d);
            const int oa = __shfl_up_sync(0xffffffffu, arg, d);
            if (lane >= d && !(m < om)) {
                m = om;
                arg = oa;
            }
        }
        if (!(m < carry_min)) {
            m = carry_min;
            arg = carry_arg;
        }
        const double five_back = (c >= 5 && valid) ? rate[pos - 5 * dir] : at_peak;
        bool stop = !valid || pos <= 2 || pos >= T - 3;
        if (!stop) stop = avail[pos + dir] == 0 || (3 * m) < v || (v < thresh && five_back < v);
        const unsigned st = __ballot_sync(0xffffffffu, stop);
        if (st) return __shfl_sync(0xffffffffu, arg, __ffs(st) - 1);
        carry_min = __shfl_sync(0xffffffffu, m, 31);
        carry_arg = __shfl_sync(0xffffffffu, arg, 31);
    }
}

__device__ inline double bd_half_up(double x) { return (double(int(x)) < x) ? x + 0.5 : x; }  // gibbs-nipt-block.cpp:1296-1303

// grid = jobs, BD_NT threads; dynamic shared memory: rate [T] f64, peak list [T] i32, availability [T] u8 when they fit
// (use_smem), else the same arrays in the job's global scratch.
__global__ void __launch_bounds__(BD_NT) k_block_define(BatchParams P, const JobDev* __restrict__ jobs, int use_smem) {
    extern __shared__ __align__(16) unsigned char bd_dyn[];
    __shared__ int s_part[BD_NT / 32];
    __shared__ int s_count;
    __shared__ double s_thresh;
    const JobDev& J = jobs[blockIdx.x];
    if (*J.underflow) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = P.T, R = J.R, n1 = T - 1;
    BlockScratch B;
    B.carve(J.blk, T);
    double* rate = use_smem ? reinterpret_cast<double*>(bd_dyn) : B.smoothed;
    int32_t* peaks = use_smem ? reinterpret_cast<int32_t*>(bd_dyn + (size_t)T * 8) : B.idx_a;
    uint8_t* avail = use_smem ? bd_dyn + (size_t)T * 12 : B.available;
    int32_t* keep = B.to_keep;  // flag per grid: a block boundary
    const int32_t* __restrict__ Lg = J.L_grid;
    const double* __restrict__ raw = B.rate2;
    if (tid == 0) {
        s_count = 0;
        s_thresh = 1.0;
    }
    // ---- 1. rate averaged over +- shuffle_bin_radius bp around the midpoint of grids i and i + 1: every inter-grid interval
    //         weighs in with the base pairs of it that fall inside the window
    for (int i = tid; i < n1; i += BD_NT) {
        const int mid = (Lg[i] + Lg[i + 1]) / 2;
        double acc = 0, bp = 0;
        int budget = P.shuffle_bin_radius, edge = mid;
        for (int j = i; j >= 0 && budget > 0; j--) {
            const int span = min(edge - Lg[j], budget);
            budget -= span;
            acc = acc + span * raw[j];
            bp += span;
            edge = Lg[j];
        }
        budget = P.shuffle_bin_radius;
        edge = mid;
        for (int j = i + 1; j < T && budget > 0; j++) {
            const int span = min(Lg[j] - edge, budget);
            budget -= span;
            acc = acc + span * raw[j - 1];
            bp += span;
            edge = Lg[j];
        }
        rate[i] = acc / bp;
    }
    for (int g = tid; g < T; g += BD_NT) {
        keep[g] = 0;
        peaks[g] = 0;
    }
    __syncthreads();
    // ---- 2. stable ranks by counting: ascending rank v = int(n1 * q) is the quantile (gibbs-nipt-block.cpp:81-85),
    //         the descending order is the peak list (ties: lower index first in both, as a stable sort leaves them)
    const int qpos = int(n1 * P.block_q);
    for (int i = tid; i < n1; i += BD_NT) {
        const double x = rate[i];
        int less = 0, greater = 0, tie_before = 0;
        for (int j = 0; j < n1; j++) {
            const double y = rate[j];
            less += (y < x) ? 1 : 0;
            greater += (y > x) ? 1 : 0;
            tie_before += (y == x && j < i) ? 1 : 0;
        }
        if (less + tie_before == qpos) s_thresh = fmin(1.0, x);
        peaks[min(greater + tie_before, n1 - 1)] = i;
    }
    __syncthreads();
    const double thresh = s_thresh;
    {
        int mine = 0;
        for (int i = tid; i < n1; i += BD_NT) {
            const uint8_t a = (thresh < rate[i]) ? 1 : 0;
            avail[i] = a;
            mine += a;
        }
        if (tid == 0) avail[n1] = 0;
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, d);
        if (lane == 0 && mine) atomicAdd(&s_count, mine);
    }
    __syncthreads();
    const int n_avail = s_count;
    // ---- 3. peaks in order of decreasing rate; a peak with both neighbours still available claims its valley
    if (warp == 0 && n_avail > 0) {
        int at = 0;
        while (at < n_avail) {
            const int q = at + lane;
            const int cand = (q < n_avail) ? min(max(peaks[q], 0), n1 - 1) : -1;
            const unsigned live = __ballot_sync(0xffffffffu, cand >= 0 && avail[cand] != 0);
            if (!live) {
                at += 32;
                continue;
            }
            const int first = __ffs(live) - 1;
            const int peak = __shfl_sync(0xffffffffu, cand, first);
            at += first + 1;
            int lo = max(peak - 1, 0), hi = min(peak + 1, n1 - 1);
            if (hi - lo == 2 && avail[lo] && avail[hi]) {  // (the peak itself is available)
                lo = bd_valley(rate, avail, peak, thresh, T, -1, lane);
                hi = bd_valley(rate, avail, peak, thresh, T, +1, lane);
            }
            __syncwarp();
            for (int j = lo + lane; j <= hi; j += 32) avail[j] = 0;
            if (lane == 0) keep[peak + 1] = 1;
            __syncwarp();
        }
        if (lane == 0) {
            keep[0] = 1;
            keep[T - 1] = 1;
        }
    }
    __syncthreads();
    // ---- 4. block id of a grid = boundaries at or before it - 1 (the last grid stays in the last block)
    int32_t* bgrid = B.blocked_grid;
    if (n_avail > 0) {
        const int n_keep = bd_scan_excl(keep, bgrid, T, s_part);
        for (int g = tid; g < T; g += BD_NT) bgrid[g] = min(bgrid[g] + keep[g] - 1, n_keep - 2);
    } else {
        for (int g = tid; g < T; g += BD_NT) bgrid[g] = 0;
    }
    __syncthreads();
    int n_blocks = bgrid[T - 1] + 1;
    // grid ranges: a block ends where the id rises (or at the last grid)
    int32_t* gs = B.grid_start;
    int32_t* ge = B.grid_end;
    int32_t* rs = B.reads_start;
    int32_t* re = B.reads_end;
    int32_t* flag = B.rmflag;
    int32_t* seg = B.idx_b;  // first read of the run of reads that currently maps to a block
    for (int g = tid; g < T; g += BD_NT) flag[g] = (g == T - 1 || bgrid[g] < bgrid[g + 1]) ? 1 : 0;
    __syncthreads();
    bd_scan_excl(flag, B.grid_where, T, s_part);  // (grid_where is scratch here; it is filled at the end)
    for (int g = tid; g < T; g += BD_NT) {
        if (flag[g]) {
            const int b = B.grid_where[g];
            ge[b] = g;
            if (g + 1 < T) gs[b + 1] = g + 1;
        }
        if (g == 0) gs[0] = 0;
    }
    for (int b = tid; b < n_blocks; b += BD_NT) {
        rs[b] = -1;
        re[b] = -1;
    }
    __syncthreads();
    // read ranges (gibbs-nipt-block.cpp:1400-1436): reads are ordered by grid, so the block id never falls.  A run of reads
    // is recorded when a LATER read (before the last one) opens a new block; the last read records its own block with the
    // start of the run that was open when it arrived.
    auto block_of_read = [&](int r) { return bgrid[J.wif0[r]]; };
    for (int r = tid; r < R - 1; r += BD_NT)
        if (r == 0 || block_of_read(r) > block_of_read(r - 1)) seg[block_of_read(r)] = r;
    __syncthreads();
    for (int r = tid + 1; r < R; r += BD_NT) {
        const int before = block_of_read(r - 1);
        if (r == R - 1) {
            const int b = block_of_read(r);
            rs[b] = seg[before];
            re[b] = r;
        } else if (block_of_read(r) > before) {
            rs[before] = seg[before];
            re[before] = r - 1;
        }
    }
    __syncthreads();
    // ---- blocks without reads (gibbs-nipt-block.cpp:1440-1530): every maximal run [s, e] of them hands its grids to the
    //      neighbours, split at the half-way point; runs touch disjoint entries, one thread per run
    for (int b = tid; b < n_blocks; b += BD_NT) flag[b] = (rs[b] == -1) ? 1 : 0;
    __syncthreads();
    for (int b = tid; b < n_blocks; b += BD_NT) {
        if (flag[b] && (b == 0 || !flag[b - 1])) {
            int s = b, e = b;
            while (e + 1 < n_blocks && flag[e + 1]) e++;
            double x = bd_half_up(0.5 * double(gs[s] + ge[e]));
            if (s == 0) {
                s = 1;
                x = 0;
            }
            if (e == n_blocks - 1) {
                e = e - 1;
                x = ge[n_blocks - 1];
            }
            gs[e + 1] = (int)x;
            ge[s - 1] = (int)(x - 1);
        }
    }
    __syncthreads();
    {
        // compaction: surviving blocks move up (through scratch, then back)
        int32_t* t0 = reinterpret_cast<int32_t*>(B.smoothed);  // [2 T] ints: the smoothed rate is no longer needed
        int32_t* t1 = t0 + T;
        int32_t* t2 = B.idx_a;
        int32_t* t3 = B.to_keep;
        int32_t* posn = B.grid_where;
        for (int b = tid; b < n_blocks; b += BD_NT) flag[b] = 1 - flag[b];
        __syncthreads();
        const int n_left = bd_scan_excl(flag, posn, n_blocks, s_part);
        if (n_left != n_blocks) {
            for (int b = tid; b < n_blocks; b += BD_NT) {
                if (flag[b]) {
                    const int o = posn[b];
                    t0[o] = rs[b];
                    t1[o] = re[b];
                    t2[o] = gs[b];
                    t3[o] = ge[b];
                }
            }
            __syncthreads();
            for (int b = tid; b < n_left; b += BD_NT) {
                rs[b] = t0[b];
                re[b] = t1[b];
                gs[b] = t2[b];
                ge[b] = t3[b];
            }
            n_blocks = n_left;
        }
        __syncthreads();
    }
    for (int g = tid; g < T; g += BD_NT) B.grid_where[g] = -1;
    __syncthreads();
    for (int b = tid; b < n_blocks; b += BD_NT) B.grid_where[ge[b]] = b;
    if (tid == 0) B.n_blocks[0] = n_blocks;
}

// emission value of read r (descriptor d, staged nowhere: global tables) for haplotype k at grid g
__device__ __forceinline__ double read_emission_global(const JobDev& J, const ReadDesc& d, int Kp, int g, int k) {
    if (d.mode == MODE_DENSE) return J.dense[(size_t)d.off * Kp + k].E;
    return J.tabs[d.off + read_pattern_global(d, J.W, Kp, g, k)].E;
}

// The block resampler proper.  grid = jobs, NT threads.  Of the reference's 6 x 3 forward vectors only nine are
// distinct (slot h run with emission label i; a permutation just picks three of them), kept in registers.
template <int NT, int EPT>
__global__ void __launch_bounds__(NT) k_block_nipt(BatchParams P, const JobDev* __restrict__ jobs, int episode) {
    __shared__ double red[2 * SW_VMAX * (NT / 32)];
    __shared__ JobDev Js;
    __shared__ int s_ns[8];
    __shared__ int s_dec[2];  // ir_chosen, do_change
    __shared__ double s_D[9];
    const int tid = threadIdx.x;
    if (tid == 0) Js = jobs[blockIdx.x];
    __syncthreads();
    const JobDev& J = Js;
    if (*J.underflow) return;
    const int K = P.K, Kp = P.Kp, T = P.T, R = J.R;
    const double ff = P.ff;
    BlockSumV<NT> bsum(red);
    BlockScratch B;
    B.carve(J.blk, T);
    const double prior = P.one_over_K, one_over_K = P.one_over_K;
    const size_t hs = (size_t)T * Kp;
    // episode-stream mode: six unused runif(R) rows precede runif_block at the episode's stream position (gibbs-nipt.cpp:3013-3018)
    const double* __restrict__ runif_block =
        J.ep_stream ? J.runif_block + (size_t)ld_cg(J.lik + LIK_EP_POS) + 6 * (size_t)R : J.runif_block + (size_t)episode * R;
    const int n_blocks = B.n_blocks[0];
    // ---- scalar state of thread 0 (the reference's logC_before / logC_after)
    double logC_before[3] = {0, 0, 0}, logC_after[3] = {0, 0, 0};
    // log c_h[g] -> scratch, then accu (two interleaved accumulators, even / odd)
    for (int i = tid; i < 3 * T; i += NT) B.logc[i] = log(ld_cg(J.c + i));
    __syncthreads();
    if (tid == 0) {
        for (int h = 0; h < 3; h++) {
            double v1 = 0, v2 = 0;
            int g;
            for (g = 0; g + 1 < T; g += 2) {
                v1 += B.logc[h * T + g];
                v2 += B.logc[h * T + g + 1];
            }
            if (g < T) v1 += B.logc[h * T + g];
            logC_after[h] = v1 + v2;
        }
    }
    bool ever_changed = false;
    double A[3][3][EPT];  // [slot h][emission label i]
#pragma unroll
    for (int h = 0; h < 3; h++)
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int e = 0; e < EPT; e++) A[h][i][e] = 0.0;
    for (int g = 0; g < T; g++) {
        // ---- Rcpp_gibbs_block_forward_one: eMatGridLocal.col(i) = eMatGrid_t{i+1}.col(g)
        double el[3][EPT];
#pragma unroll
        for (int i = 0; i < 3; i++) Col<NT, EPT>::load(el[i], J.eG + i * hs + (size_t)g * Kp, K, 0.0);
        if (g + 1 < T) {
            // pull the next grid's eMatGrid columns towards L2 (one 128-byte line per 16 doubles): the walk is serial over the grids
            const int lines = (Kp + 15) >> 4;
            for (int l = tid; l < 3 * lines; l += NT) {
                const int i = l / lines, q = l - i * lines;
                prefetch_l2(J.eG + i * hs + (size_t)(g + 1) * Kp + (q << 4));
            }
        }
        double t0 = 0, jump = 0;
        if (g > 0) {
            t0 = J.tm[2 * (g - 1)];
            jump = J.tm[2 * (g - 1) + 1] * one_over_K;
        }
#pragma unroll
        for (int h = 0; h < 3; h++) {
            double sv[3];
#pragma unroll
            for (int i = 0; i < 3; i++) {
#pragma unroll
                for (int e = 0; e < EPT; e++) {
                    const bool in = tid + e * NT < K;
                    double v;
                    if (g == 0)
                        v = prior * el[i][e];
                    else
                        v = el[i][e] * (t0 * A[h][i][e] + jump);
                    A[h][i][e] = in ? v : 0.0;
                }
                sv[i] = Col<NT, EPT>::sum(A[h][i]);
            }
            bsum.run(sv);
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const double d = 1 / sv[i];
                if (tid == 0) B.lcs[(size_t)g * 9 + h * 3 + i] = log(d);
#pragma unroll
                for (int e = 0; e < EPT; e++) A[h][i][e] = d * A[h][i][e];
            }
        }
        const int iBlock = B.grid_where[g];
        if (iBlock > -1) {
            const int gs = B.grid_start[iBlock], ge = B.grid_end[iBlock];
            const int rs0 = B.reads_start[iBlock], re0 = B.reads_end[iBlock];
            // ---- Rcpp_consider_block_relabelling: D[h][i] = sum_k alphaStore-vector (h, i) * beta_h[:, g]
            if (tid < 8) s_ns[tid] = 0;
#pragma unroll
            for (int h = 0; h < 3; h++) {
                double bt[EPT];
                Col<NT, EPT>::load(bt, J.beta + h * hs + (size_t)g * Kp, K, 0.0);
                double sv[3];
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    double s = 0;
#pragma unroll
                    for (int e = 0; e < EPT; e++) s += A[h][i][e] * bt[e];
                    sv[i] = s;
                }
                bsum.run(sv);
                if (tid == 0) {
                    s_D[h * 3 + 0] = sv[0];
                    s_D[h * 3 + 1] = sv[1];
                    s_D[h * 3 + 2] = sv[2];
                }
            }
            __syncthreads();
            for (int r = rs0 + tid; r <= re0; r += NT) atomicAdd(&s_ns[J.Hclass[r]], 1);
            __syncthreads();
            if (tid == 0) {
                double choice_log_probs[6];
                for (int ir = 0; ir < 6; ir++) {
                    double Psum = 0;
                    for (int i = 0; i < 3; i++) {
                        // slot i of permutation ir carries the vector run with emission label j, RR[ir][j] == i + 1
                        int j = 0;
                        for (int q = 0; q < 3; q++)
                            if (c_RR[ir][q] == i + 1) j = q;
                        double logC_inside = 0;
                        for (int g2 = gs; g2 <= ge; g2++) logC_inside += B.lcs[(size_t)g2 * 9 + i * 3 + j];
                        const double pm = log(s_D[i * 3 + j]) + -logC_before[i] + -logC_inside + -logC_after[i];
                        Psum += pm;
                    }
                    // rcpp_calculate_block_read_label_probabilities_using_H_class (:251-279)
                    const int* rr = c_RR[ir];
                    const double lh = 0 + s_ns[rr[0]] * P.lhc[0] + s_ns[rr[1]] * P.lhc[1] + s_ns[rr[2]] * P.lhc[2] + s_ns[7 - rr[2]] * P.lhc[3] +
                                      s_ns[7 - rr[1]] * P.lhc[4] + s_ns[7 - rr[0]] * P.lhc[5];
                    choice_log_probs[ir] = lh + Psum;
                }
                double mx = choice_log_probs[0];
                for (int ir = 1; ir < 6; ir++)
                    if (choice_log_probs[ir] > mx) mx = choice_log_probs[ir];
                const double a = -mx;
                double choice_probs[6];
                for (int ir = 0; ir < 6; ir++) {
                    double v = choice_log_probs[ir] + a;
                    if (v < (-100)) v = -100;
                    choice_probs[ir] = exp(v);
                }
                if (ff == 0) {
                    choice_probs[1] = 0;
                    choice_probs[3] = 0;
                    choice_probs[4] = 0;
                    choice_probs[5] = 0;
                }
                double ssum = 0;
                for (int ir = 0; ir < 6; ir++) ssum += choice_probs[ir];
                const double d = (1 / ssum);
                for (int ir = 0; ir < 6; ir++) choice_probs[ir] *= d;
                const double chance = runif_block[iBlock];
                double cum[6];
                cum[0] = choice_probs[0];
                for (int ir = 1; ir < 6; ir++) cum[ir] = 0 + choice_probs[ir] + cum[ir - 1];
                int ir_chosen = 0;
                for (int ir = 5; ir >= 0; ir--)
                    if (chance < cum[ir]) ir_chosen = ir;
                s_dec[0] = ir_chosen;
                s_dec[1] = (ever_changed || ir_chosen != 0) ? 1 : 0;
            }
            __syncthreads();
            const int ir_chosen = s_dec[0];
            const bool do_change = s_dec[1] != 0;
            int swap8[8];
            swap8[0] = 0;
            swap8[1] = c_RX[ir_chosen][0];
            swap8[2] = c_RX[ir_chosen][1];
            swap8[3] = c_RX[ir_chosen][2];
            swap8[4] = 7 - c_RX[ir_chosen][2];
            swap8[5] = 7 - c_RX[ir_chosen][1];
            swap8[6] = 7 - c_RX[ir_chosen][0];
            swap8[7] = 7;
            if (do_change) {
                ever_changed = true;
                // ---- rewrite eMatGrid / alpha / c of the block with the permuted labels (:838-925)
                double ap[3][EPT];
                if (gs > 0) {
#pragma unroll
                    for (int h = 0; h < 3; h++) Col<NT, EPT>::load(ap[h], J.alpha + h * hs + (size_t)(gs - 1) * Kp, K, 0.0);
                }
                for (int g2 = gs; g2 <= ge; g2++) {
                    double eg[3][EPT];
#pragma unroll
                    for (int h = 0; h < 3; h++)
#pragma unroll
                        for (int e = 0; e < EPT; e++) eg[h][e] = 1.0;
                    const int r0 = J.rs[g2], r1 = J.rs[g2 + 1];
                    // the allele words of this thread's haplotypes around grid g2 are loaded once (coalesced) and shared by the
                    // grid's reads: per (read, haplotype) a shift, a mask and one table load (every branch is uniform over the CTA)
                    uint32_t wm[EPT], w0[EPT], wp[EPT];
                    if (r1 > r0) {
#pragma unroll
                        for (int e = 0; e < EPT; e++) {
                            const int k = tid + e * NT;
                            const bool live = k < K;
                            wm[e] = (live && g2 > 0) ? J.W[(size_t)(g2 - 1) * Kp + k] : 0u;
                            w0[e] = live ? J.W[(size_t)g2 * Kp + k] : 0u;
                            wp[e] = (live && g2 + 1 < T) ? J.W[(size_t)(g2 + 1) * Kp + k] : 0u;
                        }
                    }
                    for (int r = r0; r < r1; r++) {
                        const ReadDesc d = J.desc[r];
                        const int h = swap8[J.H[r]] - 1;
                        double E[EPT];
                        const bool crossing = d.b0 + d.nb > 32;
                        if (d.mode == MODE_RUN && !(crossing && d.g0rel > 0)) {
                            const uint32_t mask = (1u << d.nb) - 1u;
                            const TabEnt* __restrict__ tb = J.tabs + d.off;
#pragma unroll
                            for (int e = 0; e < EPT; e++) {
                                const uint32_t lo = d.g0rel < 0 ? wm[e] : (d.g0rel == 0 ? w0[e] : wp[e]);
                                const uint32_t hi = d.g0rel < 0 ? w0[e] : wp[e];  // (only the crossing bits survive the mask)
                                E[e] = tb[__funnelshift_r(lo, hi, d.b0) & mask].E;
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < EPT; e++) {
                                const int k = tid + e * NT;
                                E[e] = (k < K) ? read_emission_global(J, d, Kp, g2, k) : 1.0;
                            }
                        }
                        if (h == 0) {
#pragma unroll
                            for (int e = 0; e < EPT; e++) eg[0][e] *= E[e];
                        } else if (h == 1) {
#pragma unroll
                            for (int e = 0; e < EPT; e++) eg[1][e] *= E[e];
                        } else {
#pragma unroll
                            for (int e = 0; e < EPT; e++) eg[2][e] *= E[e];
                        }
                    }
                    double sv[3];
                    double t0b = 0, t1b = 0;
                    if (g2 > 0) {
                        t0b = J.tm[2 * (g2 - 1)];
                        t1b = J.tm[2 * (g2 - 1) + 1];
                    }
#pragma unroll
                    for (int h = 0; h < 3; h++) {
                        Col<NT, EPT>::store(eg[h], J.eG + h * hs + (size_t)g2 * Kp, K);
#pragma unroll
                        for (int e = 0; e < EPT; e++) {
                            const bool in = tid + e * NT < K;
                            double v;
                            if (g2 == 0)
                                v = prior * eg[h][e];
                            else
                                v = eg[h][e] * (t0b * ap[h][e] + t1b * prior);
                            ap[h][e] = in ? v : 0.0;
                        }
                        sv[h] = Col<NT, EPT>::sum(ap[h]);
                    }
                    bsum.run(sv);
#pragma unroll
                    for (int h = 0; h < 3; h++) {
                        const double cc = 1 / sv[h];
#pragma unroll
                        for (int e = 0; e < EPT; e++) ap[h][e] *= cc;
                        Col<NT, EPT>::store(ap[h], J.alpha + h * hs + (size_t)g2 * Kp, K);
                        if (tid == 0) J.c[h * T + g2] = cc;
                    }
                }
                __syncthreads();
                for (int r = rs0 + tid; r <= re0; r += NT) {
                    J.Hclass[r] = swap8[J.Hclass[r]];
                    J.H[r] = swap8[J.H[r]];
                }
                __syncthreads();
            }
            // ---- Rcpp_reset_local_variables: every permutation restarts from the (possibly rewritten) alpha of grid g
            if (iBlock + 1 < n_blocks) {
#pragma unroll
                for (int h = 0; h < 3; h++) {
                    double av[EPT];
                    Col<NT, EPT>::load(av, J.alpha + h * hs + (size_t)g * Kp, K, 0.0);
#pragma unroll
                    for (int i = 0; i < 3; i++)
#pragma unroll
                        for (int e = 0; e < EPT; e++) A[h][i][e] = av[e];
                }
                if (tid == 0) {
                    for (int h = 0; h < 3; h++) {
                        const double lc = log(J.c[h * T + g]);
                        for (int i = 0; i < 3; i++) B.lcs[(size_t)g * 9 + h * 3 + i] = lc;
                    }
                }
            }
            if (tid == 0) {
                for (int g2 = gs; g2 <= ge; g2++)
                    for (int h = 0; h < 3; h++) logC_before[h] += log(J.c[h * T + g2]);
            }
        }
        if (tid == 0) {
            for (int h = 0; h < 3; h++) logC_after[h] -= log(J.c[h * T + g]);
        }
        __syncthreads();
    }
}

// Rcpp::sample(1:3, 1, false, probs) for three outcomes: normalise, sort descending carrying the index (bubble order),
// cumulate, first j with rU <= p[j]
__device__ inline int sample_1_of_3(const double* probs_in, double rU) {
    double p[3];
    int perm[3] = {1, 2, 3};
    const double s = probs_in[0] + probs_in[1] + probs_in[2];
    for (int i = 0; i < 3; i++) p[i] = probs_in[i] / s;
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2 - i; j++)
            if (p[j] < p[j + 1]) {
                const double tp = p[j];
                p[j] = p[j + 1];
                p[j + 1] = tp;
                const int ti = perm[j];
                perm[j] = perm[j + 1];
                perm[j + 1] = ti;
            }
    for (int i = 1; i < 3; i++) p[i] += p[i - 1];
    int j;
    for (j = 0; j < 2; j++)
        if (rU <= p[j]) break;
    return perm[j];
}

// rcpp_sample_H_using_H_class (gibbs-nipt-block.cpp:213-246): labels redrawn from the H_class of every read; a read of
// class 0 / 4 / 5 / 6 / 7 consumes the next uniform.  grid = jobs, 256 threads (chunked prefix count).
__global__ void __launch_bounds__(256) k_sample_H(BatchParams P, const JobDev* __restrict__ jobs, int episode) {
    __shared__ int s_cnt[257];
    const JobDev& J = jobs[blockIdx.x];
    if (*J.underflow) return;
    const int R = J.R, tid = threadIdx.x;
    const double ff = P.ff;
    // episode-stream mode: the H_class draws follow the episode's 8 R block-Gibbs uniforms; the position then moves past them
    const size_t ep_pos = J.ep_stream ? (size_t)ld_cg(J.lik + LIK_EP_POS) : 0;
    const double* __restrict__ ru = J.ep_stream ? J.runif_block + ep_pos + 8 * (size_t)R : J.runif_H_class + (size_t)episode * R;
    const int per = (R + 255) / 256;
    const int a = min(tid * per, R), b = min(a + per, R);
    int n = 0;
    for (int r = a; r < b; r++) {
        const int hc = J.Hclass[r];
        n += (hc == 0 || hc >= 4) ? 1 : 0;
    }
    s_cnt[tid + 1] = n;
    if (tid == 0) s_cnt[0] = 0;
    __syncthreads();
    if (tid == 0)
        for (int i = 1; i <= 256; i++) s_cnt[i] += s_cnt[i - 1];
    __syncthreads();
    int used = s_cnt[tid];
    const double probs07[3] = {0.5, 0.5 - ff * 0.5, ff * 0.5};
    const double probs4[3] = {0.5, 0.5 - 0.5 * ff, 0};
    const double probs5[3] = {0.5, 0, 0.5 * ff};
    const double probs6[3] = {0, 0.5 - ff * 0.5, ff * 0.5};
    for (int r = a; r < b; r++) {
        const int hc = J.Hclass[r];
        int v;
        if (hc == 0 || hc == 7)
            v = sample_1_of_3(probs07, ru[used++]);
        else if (hc == 4)
            v = sample_1_of_3(probs4, ru[used++]);
        else if (hc == 5)
            v = sample_1_of_3(probs5, ru[used++]);
        else if (hc == 6)
            v = sample_1_of_3(probs6, ru[used++]);
        else
            v = hc;  // classes 1, 2, 3 are the label itself
        J.H[r] = v;
    }
    if (J.ep_stream && tid == 0) J.lik[LIK_EP_POS] = (double)(ep_pos + 8 * (size_t)R + (size_t)s_cnt[256]);
}

// beta[:, T-1] = c[T-1], then Rcpp_run_backward_haploid_QUILT_faster (copied-from-stitch.cpp:417-440).  grid = (jobs, NH)
template <int NT, int EPT>
__global__ void __launch_bounds__(NT) k_bwd_fast(BatchParams P, const JobDev* __restrict__ jobs) {
    __shared__ double red[2 * SW_VMAX * (NT / 32)];
    const JobDev& J = jobs[blockIdx.x];
    if (*J.underflow) return;
    const int h = blockIdx.y, tid = threadIdx.x;
    const int K = P.K, Kp = P.Kp, T = P.T;
    BlockSumV<NT> bsum(red);
    const double one_over_K = P.one_over_K;
    const double* eG = J.eG + (size_t)h * T * Kp;
    double* beta = J.beta + (size_t)h * T * Kp;
    const double* c = J.c + h * T;
    double b[EPT], e[EPT];
    const double clast = ld_cg(c + T - 1);
#pragma unroll
    for (int i = 0; i < EPT; i++) b[i] = (tid + i * NT < K) ? clast : 0.0;
    Col<NT, EPT>::store(b, beta + (size_t)(T - 1) * Kp, K);
    for (int g = T - 2; g >= 0; g--) {
        const bool has1 = J.rs[g + 2] > J.rs[g + 1];
        const double cg = ld_cg(c + g);
        const double t0 = J.tm[2 * g], t1 = J.tm[2 * g + 1];
        if (has1) {
            Col<NT, EPT>::load(e, eG + (size_t)(g + 1) * Kp, K, 0.0);
#pragma unroll
            for (int i = 0; i < EPT; i++) b[i] = e[i] * b[i];
        }
        double sv[1] = {Col<NT, EPT>::sum(b)};
        bsum.run(sv);
        const double x = t1 * sv[0] * one_over_K;
#pragma unroll
        for (int i = 0; i < EPT; i++) b[i] = (tid + i * NT < K) ? cg * (x + t0 * b[i]) : 0.0;
        Col<NT, EPT>::store(b, beta + (size_t)g * Kp, K);
    }
}

}  // namespace qb
