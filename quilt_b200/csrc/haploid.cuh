// haploid.cuh — the full-panel haploid Li-Stephens pass: one haplotype's genotype likelihoods against ALL panel haplotypes.
//
// Reference: Rcpp_haploid_dosage_versus_refs (QUILT/src/reference-single.cpp:2189-2413) with the production settings of its
// caller (QUILT/R/functions.R:2034-2070): per-grid emission table eMatDH (Rcpp_build_eMatDH :272-329), forward "version 3"
// (:878-1131: lazy normalisation, emissions divided by their per-grid maximum, special haplotypes recomputed from their
// bits), backward "version 3" (:1781-2179: beta, gamma, dosage through per-symbol gamma sums, best matches at the thinned
// grids via Rcpp_get_top_K_or_more_matches_while_building_gamma_eigen :196-266).
//
//  k_hap_eMatDH   grid (T, passes) x 256: the (nMaxDH + 1)-entry emission column of every grid, its max / min, "grid has a
//                 variant" flag (any genotype likelihood != 1)
//  k_hap_fb       grid (passes) x NT: forward then backward of one pass in one CTA; the K_full states are spread over the
//                 CTA (k = tid + i NT, EPT per thread, alpha / beta columns in registers), the u8 symbol column of the grid
//                 is the only panel data streamed (coalesced), the emission column lives in shared memory, alphaHat_t goes
//                 to HBM once and comes back once.  Sums over K use a fixed shuffle / shared-memory tree (the reference adds
//                 sequentially: results agree to ~1e-16 relative, dosages to ~1e-13).
#pragma once

#include "device_common.cuh"
#include "prep.cuh"
#include "types.h"

namespace qb {

constexpr int HAP_NT = 512;
constexpr int HAP_MAXTOP = 16;
constexpr int HAP_LISTCAP = 256;

struct HapParams {
    int32_t K, T, nSNPs, nMaxDH, n_thin, K_top, best_cap;
    uint32_t flags;
    double thr, ref_error;
};
struct HapJob {
    const double* gl;        // [nSNPs][2]
    const double* tm;        // [T - 1][2]
    const int32_t* cols;     // [T]
    double* eMatDH;          // [T][nMaxDH + 1]
    double* emax;            // [T]
    double* cmin;            // [T]
    uint8_t* hasvar;         // [T]
    double* alpha;           // [T][K]
    double* beta;            // [T][K] or null
    double* gamma;           // [T][K] or null
    double* c;               // [T]
    double* dosage;          // [nSNPs]
    int32_t* best;           // [n_thin][best_cap]
    double* best_val;        // [n_thin][best_cap]
    int32_t* best_cnt;       // [n_thin]
};

// emission of one 32-SNP word: prod_b (bit ? dR eps + dA (1 - eps) : dR (1 - eps) + dA eps), factors applied in SNP order
__device__ __forceinline__ double hap_word_prob(uint32_t w, const double* fA, const double* fR, int nloc) {
    double prob = 1;
    for (int b = 0; b < nloc; b++) prob *= ((w >> b) & 1u) ? fA[b] : fR[b];
    return prob;
}

__global__ void __launch_bounds__(256) k_hap_eMatDH(HapParams P, const HapJob* __restrict__ jobs, const int32_t* __restrict__ distinctHapsB) {
    const HapJob& J = jobs[blockIdx.y];
    const int g = blockIdx.x, tid = threadIdx.x, NM1 = P.nMaxDH + 1;
    __shared__ double fA[32], fR[32];
    __shared__ double red[2][8];
    __shared__ int s_var;
    const int s0 = 32 * g, nloc = min(32, P.nSNPs - s0);
    if (tid == 0) s_var = 0;
    __syncthreads();
    if (tid < nloc) {
        const double dR = J.gl[2 * (size_t)(s0 + tid)], dA = J.gl[2 * (size_t)(s0 + tid) + 1];
        const double ome = 1 - P.ref_error;
        fA[tid] = dR * P.ref_error + dA * ome;
        fR[tid] = dR * ome + dA * P.ref_error;
        if (dR != 1 || dA != 1) s_var = 1;
    }
    __syncthreads();
    // threads beyond nMaxDH stand for the pre-filled row 0 (= 1) in the minimum only: row 0 is then REPLACED by that minimum
    // (Rcpp_build_eMatDH :324-326), so the later maximum of the column is the maximum over the real rows
    double mx = 0.0, mn = 1.0;
    if (tid < P.nMaxDH) {
        const double v = hap_word_prob((uint32_t)distinctHapsB[(size_t)g * P.nMaxDH + tid], fA, fR, nloc);
        J.eMatDH[(size_t)g * NM1 + 1 + tid] = v;
        mx = v;
        mn = v;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, d));
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, d));
    }
    if ((tid & 31) == 0) {
        red[0][tid >> 5] = mx;
        red[1][tid >> 5] = mn;
    }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < 8; w++) {
            mx = fmax(mx, red[0][w]);
            mn = fmin(mn, red[1][w]);
        }
        // row 0 = min of the column with row 0 still 1 (Rcpp_build_eMatDH :324-326); it takes part in the later max / min
        J.eMatDH[(size_t)g * NM1] = mn;
        J.emax[g] = mx;  // (row 0 = the minimum never exceeds it)
        J.cmin[g] = mn;
        J.hasvar[g] = (uint8_t)s_var;
    }
}

// block-wide sum and minimum, every thread gets both (two barriers)
template <int NT>
__device__ __forceinline__ void hap_block_sum_min(double& s, double& m, double* scr /*[2][NT / 32]*/) {
    constexpr int NW = NT / 32;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, d);
        m = fmin(m, __shfl_xor_sync(0xffffffffu, m, d));
    }
    __syncthreads();  // (scr may still be read from the previous call)
    if ((threadIdx.x & 31) == 0) {
        scr[threadIdx.x >> 5] = s;
        scr[NW + (threadIdx.x >> 5)] = m;
    }
    __syncthreads();
    double ss = 0, mm = scr[NW];
#pragma unroll
    for (int w = 0; w < NW; w++) {
        ss += scr[w];
        mm = fmin(mm, scr[NW + w]);
    }
    s = ss;
    m = mm;
}

template <int EPT>
__global__ void __launch_bounds__(HAP_NT) k_hap_fb(HapParams P, const HapJob* __restrict__ jobs, PanelDev PD) {
    constexpr int NT = HAP_NT, NW = NT / 32;
    extern __shared__ __align__(16) unsigned char hsm[];
    const HapJob& J = jobs[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = P.K, T = P.T, NM1 = P.nMaxDH + 1;
    double* scol = reinterpret_cast<double*>(hsm);                 // [NM1] emission column of the grid (scaled)
    double* fA = scol + ((NM1 + 1) & ~1);                           // [32]
    double* fR = fA + 32;                                           // [32]
    double* scr = fR + 32;                                          // [2][NW]
    const int MGS = max(NM1, 32);                                   // row stride of mgw (it is reused for [NW][32] dosage partials)
    double* mgw = scr + 2 * NW;                                     // [NW][MGS] per-warp symbol sums
    double* dpart = mgw + (size_t)NW * MGS;                         // [NW][32] dosage partials
    unsigned long long* wtop = reinterpret_cast<unsigned long long*>(dpart + NW * 32);  // [NW][HAP_MAXTOP]
    uint32_t* swords = reinterpret_cast<uint32_t*>(wtop + NW * HAP_MAXTOP);              // [NM1] the grid's table words (row 0 unused)
    double* sk = reinterpret_cast<double*>(swords + ((NM1 + 1) & ~1));                   // [K] per-haplotype scratch: special emissions / gamma
    __shared__ unsigned long long s_thr;
    __shared__ int s_cnt;
    __shared__ int s_lk[HAP_LISTCAP];
    __shared__ double s_lv[HAP_LISTCAP];
    const double ome = 1 - P.ref_error, eps = P.ref_error;
    const double double_K = (double)K, one_over_K = 1 / (double)K;
    const uint8_t* __restrict__ hm = PD.hapMatcherR;
    const bool want_dosage = (P.flags & QUILT_HF_RETURN_DOSAGE) != 0, want_best = (P.flags & QUILT_HF_GET_BEST_HAPS) != 0;

    // stage the per-SNP factors of grid g (and the scaled emission column when asked)
    auto stage_grid = [&](int g, bool with_col, double r, bool zero_row0) {
        __syncthreads();
        const int s0 = 32 * g, nloc = min(32, P.nSNPs - s0);
        if (tid < nloc) {
            const double dR = J.gl[2 * (size_t)(s0 + tid)], dA = J.gl[2 * (size_t)(s0 + tid) + 1];
            fA[tid] = dR * eps + dA * ome;
            fR[tid] = dR * ome + dA * eps;
        }
        if (with_col)
            for (int i = tid; i < NM1; i += NT) {
                double v = J.eMatDH[(size_t)g * NM1 + i];
                if (r != 1.0) v *= r;  // eMatDH_col *= (1 / emission_max), only when the maximum is below 1
                scol[i] = (i == 0 && zero_row0) ? 0.0 : v;
            }
        __syncthreads();
        if (with_col) {
            // special haplotypes (symbol 0): emission from their own 32-SNP word.  The grid's rows of eMatDH_special_matrix are
            // walked directly (they are sorted by haplotype, so this is what the reference's binary search returns — including its
            // quirk that a group of ONE row yields the word 0, gibbs-small.cpp:69-105); the result waits in sk[k] for the owner.
            const int s1 = __ldg(PD.helper + g), e1 = __ldg(PD.helper + PD.Tc + g);
            const int n_sp = (s1 > 0 && e1 >= s1) ? e1 - s1 + 1 : 0;
            if (n_sp > 0) {
                for (int q = tid; q < n_sp; q += NT) {
                    const int k = __ldg(PD.special + (s1 - 1 + q));
                    const uint32_t w = (n_sp == 1) ? 0u : (uint32_t)__ldg(PD.special + PD.n_special + (s1 - 1 + q));
                    sk[k] = hap_word_prob(w, fA, fR, nloc);
                }
                __syncthreads();
            }
        }
    };
    auto special_prob = [&](int k, int g) -> double { return sk[k]; };

    // ================================================================= forward
    double a[EPT];
    {
        // grid 0 (:2307-2352): raw table values, special haplotypes from their bits, prior 1 / K
        stage_grid(0, true, 1.0, false);
        double s = 0, mdummy = 1;
#pragma unroll
        for (int i = 0; i < EPT; i++) {
            const int k = tid + i * NT;
            double v = 0;
            if (k < K) {
                const int dh = hm[k];
                const double prob = dh > 0 ? scol[dh] : special_prob(k, 0);
                v = prob * one_over_K;
            }
            a[i] = v;
            s += v;
        }
        hap_block_sum_min<NT>(s, mdummy, scr);
        const double c0 = 1 / s;
        if (tid == 0) J.c[0] = c0;
#pragma unroll
        for (int i = 0; i < EPT; i++) {
            const int k = tid + i * NT;
            a[i] = a[i] * c0;
            if (k < K) J.alpha[k] = a[i];
        }
    }
    double prev_sum = 1, running_min = 1;
    for (int g = 1; g < T; g++) {
        double cg = 1;
        const double jump_prob = J.tm[2 * (size_t)(g - 1) + 1] / double_K;
        const double jpp = jump_prob * prev_sum;  // always_normalize = FALSE
        const double njp = J.tm[2 * (size_t)(g - 1)];
        const double jd = jpp / njp;
        const bool hasvar = (g == 1) || J.hasvar[g];
        double run_total;
        if (hasvar) {
            const double emax = J.emax[g];
            const double r = (emax < 1) ? (1 / emax) : 1.0;
            double min_e = (emax < 1) ? J.cmin[g] * r : J.cmin[g];
            const double rs = 1 / emax;  // specials: prob *= (1 / emission_max)
            stage_grid(g, true, r, true);
            const uint8_t* col = hm + (size_t)g * K;
            double s = 0, m = 1e300;
#pragma unroll
            for (int i = 0; i < EPT; i++) {
                const int k = tid + i * NT;
                double v = 0;
                if (k < K) {
                    const int dh = col[k];
                    if (dh > 0) {
                        v = (jd + a[i]) * scol[dh];
                    } else {
                        double prob = special_prob(k, g);
                        prob *= rs;
                        v = (jd + a[i]) * prob;
                        m = fmin(m, prob);
                    }
                }
                a[i] = v;
                s += v;
            }
            hap_block_sum_min<NT>(s, m, scr);
            run_total = s;
            if (m < min_e) min_e = m;
            running_min *= min_e;
        } else {
#pragma unroll
            for (int i = 0; i < EPT; i++) a[i] = (tid + i * NT < K) ? jd + a[i] : 0.0;
            run_total = prev_sum / njp;
        }
        cg /= njp;
        if (running_min < P.thr || g == T - 1) {
            const double x = 1 / run_total;
#pragma unroll
            for (int i = 0; i < EPT; i++) a[i] *= x;
            cg /= run_total;
            run_total = 1;
            running_min = 1;
        }
        prev_sum = run_total;
        if (tid == 0) J.c[g] = cg;
#pragma unroll
        for (int i = 0; i < EPT; i++) {
            const int k = tid + i * NT;
            if (k < K) st_stream(J.alpha + (size_t)g * K + k, a[i]);
        }
    }
    // c is read back below: make the forward's writes visible to the whole CTA
    __threadfence_block();
    __syncthreads();

    // ================================================================= backward
    double b[EPT];
    double njp = 1, B_prev_star = 1;
#pragma unroll
    for (int i = 0; i < EPT; i++) b[i] = (tid + i * NT < K) ? 1 / njp : 0.0;
    B_prev_star = K * J.c[T - 1] * njp;
    for (int g = T - 1; g >= 0; g--) {
        if (g < T - 1) {
            const double jump_prob = J.tm[2 * (size_t)g + 1] / double_K;
            njp = J.tm[2 * (size_t)g];
            if (J.hasvar[g + 1]) {
                const double emax = J.emax[g + 1];
                const double r = (emax < 1) ? (1 / emax) : 1.0;
                const double rs = 1 / emax;
                stage_grid(g + 1, true, r, true);
                const uint8_t* col = hm + (size_t)(g + 1) * K;
                double s = 0, mdummy = 1;
#pragma unroll
                for (int i = 0; i < EPT; i++) {
                    const int k = tid + i * NT;
                    double v = 0;
                    if (k < K) {
                        const int dh = col[k];
                        if (dh > 0) {
                            v = b[i] * scol[dh];
                        } else {
                            double prob = special_prob(k, g + 1);
                            prob *= rs;
                            v = b[i] * prob;
                        }
                    }
                    b[i] = v;
                    s += v;
                }
                hap_block_sum_min<NT>(s, mdummy, scr);
                const double val = jump_prob / njp * s;
#pragma unroll
                for (int i = 0; i < EPT; i++) b[i] = (tid + i * NT < K) ? b[i] + val : 0.0;
                B_prev_star = J.c[g] * s;
            } else {
                const double val = jump_prob / njp * B_prev_star;
#pragma unroll
                for (int i = 0; i < EPT; i++) b[i] = (tid + i * NT < K) ? b[i] + val : 0.0;
                B_prev_star = J.c[g] * B_prev_star;
            }
        }
        // gamma (up to not_jump_prob): alpha of this grid comes back from HBM once
        double gm[EPT];
#pragma unroll
        for (int i = 0; i < EPT; i++) {
            const int k = tid + i * NT;
            gm[i] = (k < K) ? ld_stream(J.alpha + (size_t)g * K + k) * b[i] : 0.0;
        }
        const int tcol = (want_best && J.cols) ? J.cols[g] : -1;
        if (tcol >= 0) {
            // haplotypes whose gamma reaches the K_top-th largest value (ties included), in haplotype order (:196-266)
            unsigned long long loc[HAP_MAXTOP];
#pragma unroll
            for (int t = 0; t < HAP_MAXTOP; t++) loc[t] = 0ull;
#pragma unroll
            for (int i = 0; i < EPT; i++) {
                if (tid + i * NT < K) {
                    unsigned long long key = (unsigned long long)__double_as_longlong(gm[i]);  // positive doubles order like their bits
                    if (key > loc[HAP_MAXTOP - 1]) {
#pragma unroll
                        for (int t = 0; t < HAP_MAXTOP; t++) {
                            if (key > loc[t]) {
                                const unsigned long long o = loc[t];
                                loc[t] = key;
                                key = o;
                            }
                        }
                    }
                }
            }
            // K_top-th largest with multiplicity: pop the maximum K_top times (warp level, then warp 0 over the warp lists)
            int head = 0;
            for (int rr = 0; rr < P.K_top; rr++) {
                unsigned long long cand = 0ull;
#pragma unroll
                for (int t = 0; t < HAP_MAXTOP; t++)
                    if (t == head) cand = loc[t];
                unsigned long long m = cand;
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) {
                    const unsigned long long o = __shfl_xor_sync(0xffffffffu, m, d);
                    m = o > m ? o : m;
                }
                // equal values may sit in several lanes: exactly one of them pops (the lowest lane holding the maximum)
                const unsigned holders = __ballot_sync(0xffffffffu, cand == m && m != 0ull);
                if (holders && lane == __ffs(holders) - 1) head++;
                if (lane == 0) wtop[warp * HAP_MAXTOP + rr] = m;
            }
            __syncthreads();
            if (warp == 0) {
                // merge NW sorted lists of K_top keys: lane w walks list w
                int hd = 0;
                unsigned long long m = 0ull;
                for (int rr = 0; rr < P.K_top; rr++) {
                    const unsigned long long cand = (lane < NW && hd < P.K_top) ? wtop[lane * HAP_MAXTOP + hd] : 0ull;
                    m = cand;
#pragma unroll
                    for (int d = 16; d >= 1; d >>= 1) {
                        const unsigned long long o = __shfl_xor_sync(0xffffffffu, m, d);
                        m = o > m ? o : m;
                    }
                    const unsigned holders = __ballot_sync(0xffffffffu, cand == m && m != 0ull);
                    if (holders && lane == __ffs(holders) - 1) hd++;
                }
                if (lane == 0) {
                    s_thr = m;  // the K_top-th largest (0 when fewer than K_top positive values exist: everything qualifies)
                    s_cnt = 0;
                }
            }
            __syncthreads();
            const unsigned long long thr = s_thr;
#pragma unroll
            for (int i = 0; i < EPT; i++) {
                const int k = tid + i * NT;
                if (k < K && (unsigned long long)__double_as_longlong(gm[i]) >= thr) {
                    const int pos = atomicAdd(&s_cnt, 1);
                    if (pos < HAP_LISTCAP) {
                        s_lk[pos] = k;
                        s_lv[pos] = gm[i] * njp;  // special_multiplication_value = not_jump_prob (:2031)
                    }
                }
            }
            __syncthreads();
            if (tid == 0) {
                const int n = min(s_cnt, HAP_LISTCAP);
                for (int x = 1; x < n; x++) {  // haplotype order
                    const int kk = s_lk[x];
                    const double vv = s_lv[x];
                    int y = x - 1;
                    while (y >= 0 && s_lk[y] > kk) {
                        s_lk[y + 1] = s_lk[y];
                        s_lv[y + 1] = s_lv[y];
                        y--;
                    }
                    s_lk[y + 1] = kk;
                    s_lv[y + 1] = vv;
                }
                J.best_cnt[tcol] = s_cnt;
                for (int x = 0; x < P.best_cap; x++) {  // unused entries: haplotype -1, value 0
                    J.best[(size_t)tcol * P.best_cap + x] = (x < n) ? s_lk[x] : -1;
                    J.best_val[(size_t)tcol * P.best_cap + x] = (x < n) ? s_lv[x] : 0.0;
                }
            }
        }
        if (want_dosage) {
            // matched_gammas(symbol) = sum of gamma over the haplotypes showing the symbol (:2083-2096) — through the panel's
            // per-grid index of haplotypes sorted by symbol (built once per panel): gamma goes to shared memory, warp w adds the
            // segments of symbols w, w + NW, ... in a fixed order; special haplotypes bit by bit (:2101-2128); then the table
            // (:2133-2139) with the grid's words staged in shared memory.
            __syncthreads();
#pragma unroll
            for (int i = 0; i < EPT; i++) {
                const int k = tid + i * NT;
                if (k < K) sk[k] = gm[i];
            }
            for (int i = tid; i < NM1; i += NT) swords[i] = (i > 0) ? (uint32_t)__ldg(PD.distinctHapsB + (size_t)g * P.nMaxDH + (i - 1)) : 0u;
            __syncthreads();
            const int nloc = min(32, P.nSNPs - 32 * g);
            const uint16_t* __restrict__ perm = PD.hap_perm + (size_t)g * K;
            const int32_t* __restrict__ soff = PD.hap_symoff + (size_t)g * (NM1 + 1);
            for (int sym = 1 + warp; sym < NM1; sym += NW) {
                const int o0 = soff[sym], o1 = soff[sym + 1];
                double acc = 0;
                for (int q = o0 + lane; q < o1; q += 32) acc += sk[perm[q]];
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
                if (lane == 0) scol[sym] = acc * njp;  // matched_gammas *= not_jump_prob
            }
            {
                // special haplotypes: segment 0 of the index is the grid's special rows in order
                const int s1 = __ldg(PD.helper + g), e1 = __ldg(PD.helper + PD.Tc + g);
                const int n_sp = (s1 > 0 && e1 >= s1) ? e1 - s1 + 1 : 0;
                double acc = 0;
                for (int q = warp; q < n_sp; q += NW) {
                    const int k = __ldg(PD.special + (s1 - 1 + q));
                    const uint32_t w = (n_sp == 1) ? 0u : (uint32_t)__ldg(PD.special + PD.n_special + (s1 - 1 + q));
                    const double gk = sk[k] * njp;
                    acc += gk * (((w >> lane) & 1u) ? ome : eps);
                }
                dpart[warp * 32 + lane] = acc;
            }
            __syncthreads();
            {
                double acc = 0;
                if (lane < nloc)
                    for (int dh = warp; dh < P.nMaxDH; dh += NW) acc += (((swords[dh + 1] >> lane) & 1u) ? ome : eps) * scol[dh + 1];
                mgw[warp * 32 + lane] = acc;
            }
            __syncthreads();
            if (tid < nloc) {
                double sd = 0;
#pragma unroll
                for (int w = 0; w < NW; w++) sd += dpart[w * 32 + tid];
#pragma unroll
                for (int w = 0; w < NW; w++) sd += mgw[w * 32 + tid];
                J.dosage[32 * g + tid] = sd;
            }
        }
        const double x = J.c[g] * njp;
#pragma unroll
        for (int i = 0; i < EPT; i++) {
            const int k = tid + i * NT;
            b[i] *= x;
            if (k < K) {
                if (J.beta) st_stream(J.beta + (size_t)g * K + k, b[i]);
                if (J.gamma) st_stream(J.gamma + (size_t)g * K + k, gm[i] * njp);
            }
        }
    }
}

__host__ inline size_t hap_smem_bytes(int nMaxDH, int K) {
    const int NM1 = nMaxDH + 1, NW = HAP_NT / 32, MGS = NM1 > 32 ? NM1 : 32;
    return (size_t)(((NM1 + 1) & ~1) + 64 + 2 * NW + (size_t)NW * MGS + NW * 32) * 8 + (size_t)NW * HAP_MAXTOP * 8 + (size_t)((NM1 + 1) & ~1) * 4 + (size_t)K * 8;
}

}  // namespace qb
