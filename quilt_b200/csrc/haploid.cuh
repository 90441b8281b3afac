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

// block-wide sum and minimum, every thread gets both; two alternating scratch sets -> one barrier per call
template <int NT>
__device__ __forceinline__ void hap_block_sum_min(double& s, double& m, double* scr /*[2][2][NT / 32]*/, int& phase) {
    constexpr int NW = NT / 32;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, d);
        m = fmin(m, __shfl_xor_sync(0xffffffffu, m, d));
    }
    double* buf = scr + phase * (2 * NW);
    phase ^= 1;
    if ((threadIdx.x & 31) == 0) {
        buf[threadIdx.x >> 5] = s;
        buf[NW + (threadIdx.x >> 5)] = m;
    }
    __syncthreads();
    double ss = 0, mm = buf[NW];
#pragma unroll
    for (int w = 0; w < NW; w++) {
        ss += buf[w];
        mm = fmin(mm, buf[NW + w]);
    }
    s = ss;
    m = mm;
}

// What a step needs from HBM besides its own alpha column — the grid's emission column, genotype-likelihood factors, symbol
// column, special-haplotype range, transition pair, flags — is loaded into registers ONE STEP AHEAD (every thread its share)
// and handed to shared memory at the top of the step that uses it: no step of the serial walk waits for a dependent global load.
template <int EPT>
struct HapStage {
    double e;            // eMatDH[g][tid]            (tid < nMaxDH + 1)
    double dR, dA;       // gl[32 g + tid][0 / 1]      (tid < 32)
    uint8_t sym[EPT];    // hapMatcherR[g][tid + i NT]
    double tm0, tm1, emax, cmin, cg;  // transition pair INTO the grid (forward) / out of it (backward), column max / min, c[g]
    int hasvar, sp0, sp1;             // "grid has a variant", first / last special row (1-based, helper)
};

template <int EPT>
__global__ void __launch_bounds__(HAP_NT) k_hap_fb(HapParams P, const HapJob* __restrict__ jobs, PanelDev PD) {
    constexpr int NT = HAP_NT, NW = NT / 32;
    extern __shared__ __align__(16) unsigned char hsm[];
    const HapJob& J = jobs[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = P.K, T = P.T, NM1 = P.nMaxDH + 1;
    double* scol = reinterpret_cast<double*>(hsm);                 // [NM1] emission column of the grid (as stored; scaled at use)
    double* fA = scol + ((NM1 + 1) & ~1);                           // [32]
    double* fR = fA + 32;                                           // [32]
    double* scr = fR + 32;                                          // [2][2][NW]
    double* msum = scr + 4 * NW;                                    // [NM1] per-symbol gamma sums
    double* dpart = msum + ((NM1 + 1) & ~1);                        // [2][NW][32] dosage partials (special haplotypes / table)
    unsigned long long* wtop = reinterpret_cast<unsigned long long*>(dpart + 2 * NW * 32);  // [NW][HAP_MAXTOP]
    uint32_t* swords = reinterpret_cast<uint32_t*>(wtop + NW * HAP_MAXTOP);                  // [NM1] the grid's table words (row 0 unused)
    int32_t* ssoff = reinterpret_cast<int32_t*>(swords + ((NM1 + 1) & ~1));                  // [NM1 + 1] segment offsets of the symbol index
    double* sk = reinterpret_cast<double*>(ssoff + ((NM1 + 2 + 1) & ~1));                    // [K] per-haplotype scratch: special emissions / gamma
    uint16_t* sperm = reinterpret_cast<uint16_t*>(sk + K);                                   // [K] haplotypes sorted by symbol (backward, dosage)
    __shared__ unsigned long long s_thr;
    __shared__ int s_cnt;
    __shared__ int s_lk[HAP_LISTCAP];
    __shared__ double s_lv[HAP_LISTCAP];
    const double ome = 1 - P.ref_error, eps = P.ref_error;
    const double double_K = (double)K, one_over_K = 1 / (double)K;
    const uint8_t* __restrict__ hm = PD.hapMatcherR;
    const bool want_dosage = (P.flags & QUILT_HF_RETURN_DOSAGE) != 0, want_best = (P.flags & QUILT_HF_GET_BEST_HAPS) != 0;
    int phase = 0;

    // loads of grid g's data (tm_g: index of the transition pair that travels with it, < 0: none)
    auto fetch = [&](HapStage<EPT>& S, int g, int tm_g) {
        S.e = (tid < NM1) ? __ldg(J.eMatDH + (size_t)g * NM1 + tid) : 0.0;
        const int snp = 32 * g + tid;
        const bool in = tid < 32 && snp < P.nSNPs;
        S.dR = in ? __ldg(J.gl + 2 * (size_t)snp) : 1.0;
        S.dA = in ? __ldg(J.gl + 2 * (size_t)snp + 1) : 1.0;
        const uint8_t* col = hm + (size_t)g * K;
#pragma unroll
        for (int i = 0; i < EPT; i++) {
            const int k = tid + i * NT;
            S.sym[i] = (k < K) ? __ldg(col + k) : (uint8_t)1;
        }
        S.tm0 = (tm_g >= 0) ? __ldg(J.tm + 2 * (size_t)tm_g) : 1.0;
        S.tm1 = (tm_g >= 0) ? __ldg(J.tm + 2 * (size_t)tm_g + 1) : 0.0;
        S.emax = __ldg(J.emax + g);
        S.cmin = __ldg(J.cmin + g);
        S.hasvar = __ldg(J.hasvar + g);
        S.sp0 = __ldg(PD.helper + g);
        S.sp1 = __ldg(PD.helper + PD.Tc + g);
        S.cg = 0.0;
    };
    // hand a fetched stage to shared memory (emission column, per-SNP factors) and compute the special haplotypes' emissions
    // from their own 32-SNP words.  The grid's rows of eMatDH_special_matrix are walked directly (they are sorted by haplotype,
    // so this is what the reference's binary search returns — including its quirk that a group of ONE row yields the word 0,
    // gibbs-small.cpp:69-105); the result waits in sk[k] for the owner.
    auto publish = [&](const HapStage<EPT>& S, int g) {
        __syncthreads();  // (every thread is done with the previous grid's column / factors / sk)
        if (tid < NM1) scol[tid] = S.e;
        if (tid < 32) {
            fA[tid] = S.dR * eps + S.dA * ome;
            fR[tid] = S.dR * ome + S.dA * eps;
        }
        __syncthreads();
        const int n_sp = (S.sp0 > 0 && S.sp1 >= S.sp0) ? S.sp1 - S.sp0 + 1 : 0;
        if (n_sp > 0) {
            const int nloc = min(32, P.nSNPs - 32 * g);
            for (int q = tid; q < n_sp; q += NT) {
                const int k = __ldg(PD.special + (S.sp0 - 1 + q));
                const uint32_t w = (n_sp == 1) ? 0u : (uint32_t)__ldg(PD.special + PD.n_special + (S.sp0 - 1 + q));
                sk[k] = hap_word_prob(w, fA, fR, nloc);
            }
            __syncthreads();
        }
    };

    // ================================================================= forward
    double a[EPT];
    HapStage<EPT> cur, nxt;
    fetch(cur, 0, -1);
    if (T > 1) fetch(nxt, 1, 0);
    {
        // grid 0 (:2307-2352): raw table values, special haplotypes from their bits, prior 1 / K
        publish(cur, 0);
        double s = 0, mdummy = 1;
#pragma unroll
        for (int i = 0; i < EPT; i++) {
            const int k = tid + i * NT;
            double v = 0;
            if (k < K) {
                const int dh = cur.sym[i];
                const double prob = dh > 0 ? scol[dh] : sk[k];
                v = prob * one_over_K;
            }
            a[i] = v;
            s += v;
        }
        hap_block_sum_min<NT>(s, mdummy, scr, phase);
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
        cur = nxt;
        if (g + 1 < T) fetch(nxt, g + 1, g);
        double cg = 1;
        const double jump_prob = cur.tm1 / double_K;
        const double jpp = jump_prob * prev_sum;  // always_normalize = FALSE
        const double njp = cur.tm0;
        const double jd = jpp / njp;
        const bool hasvar = (g == 1) || cur.hasvar;
        double run_total;
        if (hasvar) {
            const double emax = cur.emax;
            const bool scale = emax < 1;
            const double r = 1 / emax;  // eMatDH_col *= (1 / emission_max) only when the maximum is below 1; specials always
            double min_e = scale ? cur.cmin * r : cur.cmin;
            publish(cur, g);
            double s = 0, m = 1e300;
#pragma unroll
            for (int i = 0; i < EPT; i++) {
                const int k = tid + i * NT;
                double v = 0;
                if (k < K) {
                    const int dh = cur.sym[i];
                    if (dh > 0) {
                        const double e = scale ? scol[dh] * r : scol[dh];
                        v = (jd + a[i]) * e;
                    } else {
                        const double prob = sk[k] * r;
                        v = (jd + a[i]) * prob;
                        m = fmin(m, prob);
                    }
                }
                a[i] = v;
                s += v;
            }
            hap_block_sum_min<NT>(s, m, scr, phase);
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
    // c and alpha are read back below (c by every thread, alpha by its writer): make the forward's writes visible to the CTA
    __threadfence_block();
    __syncthreads();

    // ================================================================= backward
    // step g uses the emission data of grid g + 1 (stage `cur`, with the transition pair g and c[g]) and, for the dosage, the
    // symbol index / table words of grid g itself (fetched one step ahead as well)
    struct DosStage {
        uint16_t perm[EPT];
        int32_t off;    // hap_symoff[g][tid]   (tid < NM1 + 1)
        uint32_t word;  // distinctHapsB[g][tid - 1]
        int sp0, sp1;
    };
    auto fetch_dos = [&](DosStage& D, int g) {
        const uint16_t* __restrict__ perm = PD.hap_perm + (size_t)g * K;
#pragma unroll
        for (int i = 0; i < EPT; i++) {
            const int q = tid + i * NT;
            D.perm[i] = (q < K) ? __ldg(perm + q) : (uint16_t)0;
        }
        D.off = (tid < NM1 + 1) ? __ldg(PD.hap_symoff + (size_t)g * (NM1 + 1) + tid) : 0;
        D.word = (tid > 0 && tid < NM1) ? (uint32_t)__ldg(PD.distinctHapsB + (size_t)g * P.nMaxDH + (tid - 1)) : 0u;
        D.sp0 = __ldg(PD.helper + g);
        D.sp1 = __ldg(PD.helper + PD.Tc + g);
    };
    double b[EPT];
    double njp = 1, B_prev_star = 1;
#pragma unroll
    for (int i = 0; i < EPT; i++) b[i] = (tid + i * NT < K) ? 1 / njp : 0.0;
    B_prev_star = K * J.c[T - 1] * njp;
    DosStage dcur, dnxt;
    if (want_dosage) fetch_dos(dnxt, T - 1);
    if (T > 1) {
        fetch(nxt, T - 1, T - 2);
        nxt.cg = J.c[T - 2];
    }
    double c_here = J.c[T - 1];  // c[g] of the step being processed
    for (int g = T - 1; g >= 0; g--) {
        double an[EPT];  // alpha of this grid comes back from HBM once (issued before the step's barriers)
#pragma unroll
        for (int i = 0; i < EPT; i++) {
            const int k = tid + i * NT;
            an[i] = (k < K) ? ld_stream(J.alpha + (size_t)g * K + k) : 0.0;
        }
        if (want_dosage) {
            dcur = dnxt;
            if (g > 0) fetch_dos(dnxt, g - 1);
        }
        if (g < T - 1) {
            cur = nxt;
            if (g > 0) {
                fetch(nxt, g, g - 1);
                nxt.cg = J.c[g - 1];
            }
            c_here = cur.cg;
            const double jump_prob = cur.tm1 / double_K;
            njp = cur.tm0;
            if (cur.hasvar) {
                const double emax = cur.emax;
                const bool scale = emax < 1;
                const double r = 1 / emax;
                publish(cur, g + 1);
                double s = 0, mdummy = 1;
#pragma unroll
                for (int i = 0; i < EPT; i++) {
                    const int k = tid + i * NT;
                    double v = 0;
                    if (k < K) {
                        const int dh = cur.sym[i];
                        if (dh > 0) {
                            const double e = scale ? scol[dh] * r : scol[dh];
                            v = b[i] * e;
                        } else {
                            const double prob = sk[k] * r;
                            v = b[i] * prob;
                        }
                    }
                    b[i] = v;
                    s += v;
                }
                hap_block_sum_min<NT>(s, mdummy, scr, phase);
                const double val = jump_prob / njp * s;
#pragma unroll
                for (int i = 0; i < EPT; i++) b[i] = (tid + i * NT < K) ? b[i] + val : 0.0;
                B_prev_star = c_here * s;
            } else {
                const double val = jump_prob / njp * B_prev_star;
#pragma unroll
                for (int i = 0; i < EPT; i++) b[i] = (tid + i * NT < K) ? b[i] + val : 0.0;
                B_prev_star = c_here * B_prev_star;
            }
        }
        // gamma (up to not_jump_prob)
        double gm[EPT];
#pragma unroll
        for (int i = 0; i < EPT; i++) gm[i] = an[i] * b[i];
        const int tcol = (want_best && J.cols) ? J.cols[g] : -1;
        if (tcol >= 0) {
            // haplotypes whose gamma reaches the K_top-th largest value (ties included), in haplotype order (:196-266)
            // per warp: pop the maximum K_top times.  A lane offers the largest of its elements not popped yet (bit mask of popped
            // elements — no sorted per-thread list), the warp takes the maximum, exactly one holder pops it.
            unsigned long long key[EPT];  // positive doubles order like their bits
#pragma unroll
            for (int i = 0; i < EPT; i++) key[i] = (tid + i * NT < K) ? (unsigned long long)__double_as_longlong(gm[i]) : 0ull;
            unsigned taken = 0u;
            for (int rr = 0; rr < P.K_top; rr++) {
                unsigned long long cand = 0ull;
                int ci = -1;
#pragma unroll
                for (int i = 0; i < EPT; i++) {
                    if (!((taken >> i) & 1u) && key[i] > cand) {
                        cand = key[i];
                        ci = i;
                    }
                }
                unsigned long long m = cand;
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) {
                    const unsigned long long o = __shfl_xor_sync(0xffffffffu, m, d);
                    m = o > m ? o : m;
                }
                // equal values may sit in several lanes / elements: exactly one of them pops (the lowest lane holding the maximum)
                const unsigned holders = __ballot_sync(0xffffffffu, cand == m && m != 0ull);
                if (holders && lane == __ffs(holders) - 1) taken |= 1u << ci;
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
            // per-grid index of haplotypes sorted by symbol (built once per panel, staged one step ahead): gamma goes to shared
            // memory, every 8-lane group adds the segment of one symbol in a fixed order; special haplotypes bit by bit
            // (:2101-2128); then the table (:2133-2139) with the grid's words staged in shared memory.
            __syncthreads();  // (sk / staging buffers are free)
#pragma unroll
            for (int i = 0; i < EPT; i++) {
                const int k = tid + i * NT;
                if (k < K) {
                    sk[k] = gm[i];
                    sperm[k] = dcur.perm[i];
                }
            }
            if (tid < NM1 + 1) ssoff[tid] = dcur.off;
            if (tid < NM1) swords[tid] = dcur.word;
            __syncthreads();
            const int nloc = min(32, P.nSNPs - 32 * g);
            {
                const int grp = tid >> 3, gl8 = tid & 7;  // NT / 8 groups of eight lanes, one symbol each per round
                for (int base = 1; base < NM1; base += NT / 8) {  // (trip count uniform over the CTA: the shuffles need whole warps)
                    const int sym = base + grp;
                    const bool live = sym < NM1;
                    const int o0 = live ? ssoff[sym] : 0, o1 = live ? ssoff[sym + 1] : 0;
                    double acc = 0;
                    for (int q = o0 + gl8; q < o1; q += 8) acc += sk[sperm[q]];
                    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
                    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                    if (live && gl8 == 0) msum[sym] = acc * njp;  // matched_gammas *= not_jump_prob
                }
            }
            {
                // special haplotypes: the grid's special rows in order
                const int n_sp = (dcur.sp0 > 0 && dcur.sp1 >= dcur.sp0) ? dcur.sp1 - dcur.sp0 + 1 : 0;
                double acc = 0;
                for (int q = warp; q < n_sp; q += NW) {
                    const int k = __ldg(PD.special + (dcur.sp0 - 1 + q));
                    const uint32_t w = (n_sp == 1) ? 0u : (uint32_t)__ldg(PD.special + PD.n_special + (dcur.sp0 - 1 + q));
                    const double gk = sk[k] * njp;
                    acc += gk * (((w >> lane) & 1u) ? ome : eps);
                }
                dpart[warp * 32 + lane] = acc;
            }
            __syncthreads();
            {
                double acc = 0;
                if (lane < nloc)
                    for (int dh = warp; dh < P.nMaxDH; dh += NW) acc += (((swords[dh + 1] >> lane) & 1u) ? ome : eps) * msum[dh + 1];
                dpart[(NW + warp) * 32 + lane] = acc;
            }
            __syncthreads();
            if (tid < nloc) {
                double sd = 0;
#pragma unroll
                for (int w = 0; w < NW; w++) sd += dpart[w * 32 + tid];
#pragma unroll
                for (int w = 0; w < NW; w++) sd += dpart[(NW + w) * 32 + tid];
                J.dosage[32 * g + tid] = sd;
            }
        }
        const double x = c_here * njp;
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
    const int NM1 = nMaxDH + 1, NW = HAP_NT / 32, NM1e = (NM1 + 1) & ~1;
    return (size_t)(NM1e + 64 + 4 * NW + NM1e + 2 * NW * 32) * 8 + (size_t)NW * HAP_MAXTOP * 8 + (size_t)NM1e * 4 + (size_t)((NM1 + 3) & ~1) * 4 + (size_t)K * 8 +
           (size_t)((K + 7) & ~7) * 2;
}

}  // namespace qb
