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

// stable merge sort of indices 0..n-1 by key (ascending, or descending when desc), single thread; result in idx
__device__ inline void stable_sort_idx(const double* key, int n, int32_t* idx, int32_t* tmp, bool desc) {
    for (int i = 0; i < n; i++) idx[i] = i;
    int32_t* src = idx;
    int32_t* dst = tmp;
    for (int w = 1; w < n; w <<= 1) {
        for (int lo = 0; lo < n; lo += 2 * w) {
            const int mid = min(lo + w, n), hi = min(lo + 2 * w, n);
            int i = lo, j = mid, o = lo;
            while (i < mid && j < hi) {
                const double a = key[src[i]], b = key[src[j]];
                const bool take_right = desc ? (b > a) : (b < a);  // stable: ties keep the left element
                dst[o++] = take_right ? src[j++] : src[i++];
            }
            while (i < mid) dst[o++] = src[i++];
            while (j < hi) dst[o++] = src[j++];
        }
        int32_t* t = src;
        src = dst;
        dst = t;
    }
    if (src != idx)
        for (int i = 0; i < n; i++) idx[i] = src[i];
}

// copied-from-stitch.cpp:522-567
__device__ inline int determine_where_to_stop(const double* smoothed_rate, const uint8_t* available, int snp_best, double thresh, int nGrids,
                                              bool is_left) {
    const int mult = is_left ? 1 : -1;
    int snp_consider = snp_best;
    double val_cur = smoothed_rate[snp_consider];
    double val_prev = smoothed_rate[snp_best];
    int snp_min = snp_consider;
    double val_min = smoothed_rate[snp_min];
    int c = 1;
    bool are_done = false;
    while (!are_done) {
        snp_consider = snp_consider + (-1) * mult;
        val_cur = smoothed_rate[snp_consider];
        if (5 <= c) val_prev = smoothed_rate[snp_consider + 5 * mult];
        c += 1;
        if (val_cur < val_min) {
            snp_min = snp_consider;
            val_min = val_cur;
        }
        if ((snp_consider <= 2) | ((nGrids - 3) <= snp_consider)) {
            are_done = true;
        } else if (available[snp_consider + (-1) * mult] == 0) {
            are_done = true;
        } else if ((3 * val_min) < val_cur) {
            are_done = true;
        } else if ((val_cur < thresh) & (val_prev < val_cur)) {
            are_done = true;
        }
    }
    return snp_min;
}

__device__ inline double ceiling_point5(double x) {
    if (double(int(x)) < x) return x + 0.5;
    return x;
}

// Block definition + "considers": small, strictly serial integer / scalar logic — one thread per job walks it exactly
// in the reference's order.  grid = jobs, 32 threads (lane 0 works).
__global__ void __launch_bounds__(32) k_block_define(BatchParams P, const JobDev* __restrict__ jobs) {
    if (threadIdx.x != 0) return;
    const JobDev& J = jobs[blockIdx.x];
    if (*J.underflow) return;
    const int nGrids = P.T, nReads = J.R;
    BlockScratch B;
    B.carve(J.blk, nGrids);
    const int32_t* L_grid = J.L_grid;
    const int shuffle_bin_radius = P.shuffle_bin_radius;
    double* smoothed_rate = B.smoothed;
    const double* sigma_rate = B.rate2;
    const int n1 = nGrids - 1;
    // ---- make_smoothed_rate (copied-from-stitch.cpp:446-518)
    for (int iGrid = 0; iGrid < n1; iGrid++) {
        const int focal_point = (L_grid[iGrid] + L_grid[iGrid + 1]) / 2;
        int iGrid_left = iGrid;
        int bp_remaining = shuffle_bin_radius;
        int bp_prev = focal_point;
        double total_bp_added = 0;
        double acc = 0;
        int bp_to_add;
        while ((0 < bp_remaining) & (0 <= iGrid_left)) {
            bp_to_add = (bp_prev - L_grid[iGrid_left]);
            if ((bp_remaining - bp_to_add) < 0) {
                bp_to_add = bp_remaining;
                bp_remaining = 0;
            } else {
                bp_remaining = bp_remaining - bp_to_add;
            }
            acc = acc + bp_to_add * sigma_rate[iGrid_left];
            total_bp_added += bp_to_add;
            bp_prev = L_grid[iGrid_left];
            iGrid_left = iGrid_left - 1;
        }
        int iGrid_right = iGrid + 1;
        bp_remaining = shuffle_bin_radius;
        bp_prev = focal_point;
        while ((0 < bp_remaining) & (iGrid_right < nGrids)) {
            bp_to_add = (L_grid[iGrid_right] - bp_prev);
            if ((bp_remaining - bp_to_add) < 0) {
                bp_to_add = bp_remaining;
                bp_remaining = 0;
            } else {
                bp_remaining = bp_remaining - bp_to_add;
            }
            acc = acc + bp_to_add * sigma_rate[iGrid_right - 1];
            total_bp_added += bp_to_add;
            bp_prev = L_grid[iGrid_right];
            iGrid_right = iGrid_right + 1;
        }
        smoothed_rate[iGrid] = acc / total_bp_added;
    }
    // ---- threshold = min(1, quantile) (gibbs-nipt-block.cpp:81-85, :386-392)
    double break_thresh = 1;
    {
        stable_sort_idx(smoothed_rate, n1, B.idx_a, B.idx_b, false);
        const int v = int(n1 * P.block_q);
        const double d = smoothed_rate[B.idx_a[v]];
        if (d < break_thresh) break_thresh = d;
    }
    uint8_t* available = B.available;
    int nAvailable = 0;
    for (int i = 0; i < n1; i++) {
        uint8_t av = 0;
        if (smoothed_rate[i] < 0.01) av = 0;
        if (break_thresh < smoothed_rate[i]) av = 1;
        available[i] = av;
        nAvailable += av;
    }
    int32_t* blocked_grid = B.blocked_grid;
    for (int i = 0; i < nGrids; i++) blocked_grid[i] = 0;
    int n_keep = 0;
    if (nAvailable > 0) {
        stable_sort_idx(smoothed_rate, n1, B.idx_a, B.idx_b, true);
        const int32_t* best2 = B.idx_a;
        int32_t* to_keep = B.to_keep;
        int kmin = 0x7fffffff, kmax = -1;
        for (int iBest = 0; iBest < nAvailable; iBest++) {
            const int snp_best = best2[iBest];
            if (available[snp_best]) {
                const int a = max(snp_best - 1, 0);
                const int b = min(snp_best + 1, nGrids - 1 - 1);
                int dd = 0;
                for (int j = a; j <= b; j++)
                    if (available[j]) dd += 1;
                if (dd == 3) {
                    const int snp_left = determine_where_to_stop(smoothed_rate, available, snp_best, break_thresh, nGrids, true);
                    const int snp_right = determine_where_to_stop(smoothed_rate, available, snp_best, break_thresh, nGrids, false);
                    for (int j = snp_left; j <= snp_right; j++) available[j] = 0;
                } else {
                    for (int j = a; j <= b; j++) available[j] = 0;
                }
                to_keep[n_keep++] = snp_best + 1;
                kmin = min(kmin, snp_best + 1);
                kmax = max(kmax, snp_best + 1);
            }
        }
        if (kmin != 0) to_keep[n_keep++] = 0;
        if (kmax != (nGrids - 1)) to_keep[n_keep++] = nGrids - 1;
        // sort ascending (insertion sort: the list is short)
        for (int i = 1; i < n_keep; i++) {
            const int v = to_keep[i];
            int j = i - 1;
            while (j >= 0 && to_keep[j] > v) {
                to_keep[j + 1] = to_keep[j];
                j--;
            }
            to_keep[j + 1] = v;
        }
        for (int i = 0; i < (n_keep - 1); i++) {
            const int a = to_keep[i], b = to_keep[i + 1];
            for (int j = a; j <= b; j++) blocked_grid[j] = i;
        }
    }
    // ---- Rcpp_make_gibbs_considers on blocked_snps[iSNP] = blocked_grid[iSNP / 32] (grid32): a block's SNPs are
    //      the SNPs of its grids, so grid_start / grid_end follow directly; the SNP-level outputs are not consumed
    int n_blocks = blocked_grid[nGrids - 1] + 1;
    int32_t* grid_start = B.grid_start;
    int32_t* grid_end = B.grid_end;
    {
        int iBlock = 0, start = 0;
        for (int g = 0; g < nGrids; g++) {
            const bool record = (g == nGrids - 1) || (blocked_grid[g] < blocked_grid[g + 1]);
            if (record) {
                grid_start[iBlock] = start;
                grid_end[iBlock] = g;
                start = g + 1;
                iBlock++;
            }
        }
        // the reference rebuilds blocked_grid from grid_start / grid_end (identical here)
    }
    int32_t* reads_start = B.reads_start;
    int32_t* reads_end = B.reads_end;
    for (int b = 0; b < n_blocks; b++) {
        reads_start[b] = -1;
        reads_end[b] = -1;
    }
    {
        const int32_t* wif0 = J.wif0;
        int previous_block_first_iRead = 0;
        int previous_block = blocked_grid[wif0[0]];
        for (int this_iRead = 1; this_iRead < nReads; this_iRead++) {
            const int this_block = blocked_grid[wif0[this_iRead]];
            if (this_iRead == (nReads - 1)) {
                reads_start[this_block] = previous_block_first_iRead;
                reads_end[this_block] = this_iRead;
            } else if (previous_block < this_block) {
                reads_start[previous_block] = previous_block_first_iRead;
                reads_end[previous_block] = this_iRead - 1;
                previous_block_first_iRead = this_iRead;
                previous_block = this_block;
            }
        }
    }
    // ---- removal of blocks without reads (gibbs-nipt-block.cpp:1440-1530)
    {
        int32_t* remove = B.rmflag;
        int n_to_remove = 0;
        for (int b = 0; b < n_blocks; b++) {
            remove[b] = (reads_start[b] == -1) ? 1 : 0;
            n_to_remove += remove[b];
        }
        if (n_to_remove > 0) {
            int32_t* w = B.idx_b;  // indices of the removed blocks
            int a = 0;
            for (int b = 0; b < n_blocks; b++)
                if (remove[b]) w[a++] = b;
            int jBefore = 0;
            bool todo = false;
            for (int jNow = 0; jNow < n_to_remove; jNow++) {
                if (jNow == (n_to_remove - 1)) {
                    todo = true;
                } else {
                    if ((w[jNow + 1] - w[jNow]) == 1) {
                        todo = false;
                        jBefore -= 1;
                    } else {
                        todo = true;
                    }
                }
                if (todo) {
                    int s1 = w[jBefore];
                    int e1 = w[jNow];
                    double x = ceiling_point5(0.5 * double(grid_start[s1] + grid_end[e1]));
                    if (s1 == 0) {
                        s1 = 1;
                        x = 0;
                    }
                    if (e1 == (n_blocks - 1)) {
                        e1 = e1 - 1;
                        x = grid_end[n_blocks - 1];
                    }
                    grid_start[e1 + 1] = (int)x;  // double -> int truncation, as the IntegerVector assignment
                    grid_end[s1 - 1] = (int)(x - 1);
                    jBefore = jNow;
                }
                jBefore += 1;
            }
            int o = 0;
            for (int b = 0; b < n_blocks; b++) {
                if (!remove[b]) {
                    reads_start[o] = reads_start[b];
                    reads_end[o] = reads_end[b];
                    grid_start[o] = grid_start[b];
                    grid_end[o] = grid_end[b];
                    o++;
                }
            }
            n_blocks = o;
        }
    }
    int32_t* grid_where = B.grid_where;
    for (int g = 0; g < nGrids; g++) grid_where[g] = -1;
    for (int b = 0; b < n_blocks; b++) grid_where[grid_end[b]] = b;
    B.n_blocks[0] = n_blocks;
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
                    for (int r = r0; r < r1; r++) {
                        const ReadDesc d = J.desc[r];
                        const int h = swap8[J.H[r]] - 1;
#pragma unroll
                        for (int e = 0; e < EPT; e++) {
                            const int k = tid + e * NT;
                            if (k < K) {
                                const double E = read_emission_global(J, d, Kp, g2, k);
                                if (h == 0)
                                    eg[0][e] *= E;
                                else if (h == 1)
                                    eg[1][e] *= E;
                                else
                                    eg[2][e] *= E;
                            }
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
