// sweep.cuh — the Gibbs sweep kernel: one CTA walks one sample's call through a full sweep.
//
// Reference path (QUILT/src): rcpp_gibbs_nipt_iterate gibbs-nipt.cpp:1756-1956, with
//   rcpp_alpha_forward_one_QUILT_faster   :671-707   (per-grid forward step, 1/K jump shortcut)
//   rcpp_reinitialize_in_iterations       :712-727   (grid 0)
//   sample_reads_in_grid                  :733-1295  (the read-label resampler)
//   Rcpp_run_backward_haploid_QUILT_faster copied-from-stitch.cpp:417-440
//   add_to_per_it_likelihoods             :1583-1621 (device part: -sum log c, label / H_class counts)
//
// Layout: the K states of a column are spread over the CTA (k = tid + i * NT, EPT elements per thread, in
// registers); alpha_{g-1}, the working copies alphaHat_m / ab_m of the current grid live in registers for
// the whole grid.  eMatGrid columns (both stages of a 2-deep ring) and the 32-SNP allele words of grids
// g-1 .. g+2 are staged in shared memory by 1-D bulk async copies (TMA engine, mbarrier completion); the
// small per-grid read metadata (descriptors, emission tables, labels, uniforms) by cp.async.  beta columns
// are streamed with L1-bypassing loads, alpha / beta / changed eMatGrid columns leave with streaming stores.
// Per-read K-long sums use a shuffle butterfly + one shared-memory exchange (device_common.cuh BlockSum),
// every thread ends with bit-identical totals and takes the label decision redundantly.
#pragma once

#include "device_common.cuh"
#include "types.h"

namespace qb {

constexpr int SW_MAXR = 48;     // reads of one grid staged in shared memory (more: read from global)
constexpr int SW_MAXTAB = 384;  // table entries of one grid staged in shared memory
constexpr int SW_VMAX = 4;

struct SweepSmemLayout {
    int off_bar, off_red, off_cnt, off_small[2], off_pat, off_W, off_eG, total;
    int small_desc, small_tab, small_U, small_H;
};
__host__ __device__ inline SweepSmemLayout sweep_smem_layout(int Kp, int NH, int NT) {
    SweepSmemLayout L;
    int o = 0;
    L.off_bar = o;
    o += 64;
    L.off_red = o;
    o += 2 * SW_VMAX * (NT / 32) * 8;
    o = (o + 127) & ~127;
    L.off_cnt = o;
    o += 128;
    L.small_desc = 0;
    L.small_tab = SW_MAXR * 32;
    L.small_U = L.small_tab + SW_MAXTAB * 16;
    L.small_H = L.small_U + SW_MAXR * 8;
    const int small_total = (L.small_H + SW_MAXR * 4 + 127) & ~127;
    L.off_small[0] = o;
    o += small_total;
    L.off_small[1] = o;
    o += small_total;
    L.off_pat = o;
    o += Kp * 2;
    L.off_W = o;
    o += 4 * Kp * 4;
    L.off_eG = o;
    o += 2 * NH * Kp * 8;
    L.total = o;
    return L;
}

template <int NT>
struct BlockSumV {
    static constexpr int NW = NT / 32;  // power of two (4, 8, 16)
    double* scratch;                    // [2][SW_VMAX][NW]
    int phase;
    __device__ __forceinline__ BlockSumV(double* s) : scratch(s), phase(0) {}
    // Every thread returns the same, bit-identical totals: xor butterfly inside each warp, one shared-memory slot per
    // warp, then every warp butterflies the NW partials again (same operation order in all warps).  Two alternating
    // scratch buffers make one __syncthreads per call enough.
    template <int V>
    __device__ __forceinline__ void run(double (&v)[V]) {
#pragma unroll
        for (int i = 0; i < V; i++) {
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], d);
        }
        if (NW == 1) return;
        double* buf = scratch + phase * (SW_VMAX * NW);
        phase ^= 1;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < V; i++) buf[i * NW + warp] = v[i];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < V; i++) {
            double s = buf[i * NW + (lane & (NW - 1))];
#pragma unroll
            for (int d = NW / 2; d >= 1; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
            v[i] = s;
        }
    }
};

// nearest canonical label-probability pattern (gibbs-nipt.cpp:1142-1165)
__device__ __forceinline__ int classify_H(const BatchParams& P, double x0, double x1, double x2) {
    double local_min = 2;
    int which = 8;
#pragma unroll
    for (int i = 0; i < 7; i++) {
        const double y = fabs(P.rlc[i][0] - x0) + fabs(P.rlc[i][1] - x1) + fabs(P.rlc[i][2] - x2);
        if (y < local_min) {
            local_min = y;
            which = i;
        }
    }
    return (local_min < P.class_sum_cutoff) ? which + 1 : 0;
}

// three label-indexed doubles kept in registers (dynamic indexing of a local array would go to local memory)
struct P3 {
    double a, b, c;
    __device__ __forceinline__ double get(int i) const { return i == 0 ? a : (i == 1 ? b : c); }
    __device__ __forceinline__ void set(int i, double v) {
        if (i == 0)
            a = v;
        else if (i == 1)
            b = v;
        else
            c = v;
    }
    __device__ __forceinline__ double prod() const { return (a * b) * c; }
};

// a / e from e and inv = RN(1 / e): q = RN(a * inv) is within 2 ulp, the fma residual r = a - q e is exact and
// RN(q + r inv) is the correctly rounded quotient (Markstein); 10x cheaper than the IEEE division sequence.
__device__ __forceinline__ double div_by(double a, double e, double inv) {
    const double q = a * inv;
    const double r = fma(-q, e, a);
    return fma(r, inv, q);
}

__device__ __forceinline__ uint32_t read_pattern_smem(const ReadDesc& d, const uint32_t* Wr, int Kp, int g, int k) {
    if (d.mode == MODE_RUN) {
        const int w0 = g + d.g0rel;
        const uint32_t lo = Wr[(w0 & 3) * Kp + k];
        const uint32_t hi = (d.b0 + d.nb > 32) ? Wr[((w0 + 1) & 3) * Kp + k] : 0u;
        return __funnelshift_r(lo, hi, d.b0) & ((1u << d.nb) - 1u);
    }
    uint32_t pat = 0;
    for (int j = 0; j < d.nb; j++) {
        const int wr = d.sel[j] >> 5, b = d.sel[j] & 31;
        pat |= ((Wr[((g + wr - 1) & 3) * Kp + k] >> b) & 1u) << j;
    }
    return pat;
}

template <int NT, int EPT, int NH>
__global__ void __launch_bounds__(NT, (NT <= 128 ? 5 : (NT <= 256 ? 2 : 1))) k_sweep(BatchParams P, const JobDev* __restrict__ jobs, int iteration) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ JobDev Js;
    const int tid = threadIdx.x;
    if (tid == 0) Js = jobs[blockIdx.x];
    __syncthreads();
    const JobDev& J = Js;
    if (*J.underflow) return;
    const int K = P.K, Kp = P.Kp, T = P.T, R = J.R;
    const SweepSmemLayout L = sweep_smem_layout(Kp, NH, NT);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.off_bar);
    BlockSumV<NT> bsum(reinterpret_cast<double*>(smem + L.off_red));
    int* cnt = reinterpret_cast<int*>(smem + L.off_cnt);
    uint32_t* Wr = reinterpret_cast<uint32_t*>(smem + L.off_W);
    uint16_t* spat = reinterpret_cast<uint16_t*>(smem + L.off_pat);  // [Kp] allele patterns of reads that leave the fast path
    double* eGs = reinterpret_cast<double*>(smem + L.off_eG);  // [2][NH][Kp]
    const bool iterative = (P.flags & QUILT_F_GIBBS_INITIALIZE_ITERATIVELY) != 0;
    const bool record = (P.flags & QUILT_F_RECORD_READ_SET) != 0;
    const double one_over_K = P.one_over_K;
    const int32_t* __restrict__ rs = J.rs;
    const double* __restrict__ U = J.runif_reads + (size_t)iteration * R;

    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_barrier_init();
    }
    if (tid < 16) cnt[tid] = 0;
    __syncthreads();
    uint32_t n_use0 = 0, n_use1 = 0;  // completed uses of each stage barrier -> wait parity

    // ---- one package = everything grid g needs: eMatGrid columns, the allele words of grid g+1 (ring of 4:
    //      grid g uses words g-1 .. g+1 while package g+1 is already filling word g+2), and the read
    //      metadata of the grid
    auto issue_pkg = [&](int g) {
        const int s = g & 1;
        const int r0 = rs[g], r1 = rs[g + 1];
        const int n_g = r1 - r0;
        if (tid == 0) {
            uint32_t bytes = 0;
            if (n_g > 0 || g == 0) bytes += NH * Kp * 8;
            if (g == 0) bytes += Kp * 4;
            if (g + 1 < T) bytes += Kp * 4;
            if (bytes > 0)
                mbar_arrive_expect_tx(&bar[s], bytes);
            else
                mbar_arrive(&bar[s]);
            if (n_g > 0 || g == 0) {
#pragma unroll
                for (int h = 0; h < NH; h++) bulk_g2s(eGs + (size_t)(s * NH + h) * Kp, J.eG + ((size_t)h * T + g) * Kp, Kp * 8, &bar[s]);
            }
            if (g == 0) bulk_g2s(Wr, J.W, Kp * 4, &bar[s]);
            if (g + 1 < T) bulk_g2s(Wr + ((g + 1) & 3) * Kp, J.W + (size_t)(g + 1) * Kp, Kp * 4, &bar[s]);
        }
        if (n_g > 0) {
            const int t0 = J.ts[g], nt = J.ts[g + 1] - t0;
            if (n_g <= SW_MAXR && nt <= SW_MAXTAB) {
                unsigned char* sm = smem + L.off_small[s];
                const unsigned char* gd = reinterpret_cast<const unsigned char*>(J.desc + r0);
                for (int i = tid; i < n_g * 2; i += NT) cp_async16(sm + L.small_desc + i * 16, gd + i * 16);
                const unsigned char* gt = reinterpret_cast<const unsigned char*>(J.tabs + t0);
                for (int i = tid; i < nt; i += NT) cp_async16(sm + L.small_tab + i * 16, gt + i * 16);
                for (int i = tid; i < n_g; i += NT) {
                    cp_async8(sm + L.small_U + i * 8, U + r0 + i);
                    cp_async4(sm + L.small_H + i * 4, J.H + r0 + i);
                }
            }
        }
        cp_async_commit();
    };
    auto wait_pkg = [&](int g) {
        const int s = g & 1;
        cp_async_wait_all();
        if (s == 0) {
            mbar_wait(&bar[0], n_use0 & 1);
            n_use0++;
        } else {
            mbar_wait(&bar[1], n_use1 & 1);
            n_use1++;
        }
        __syncthreads();
    };

    double am[NH][EPT];  // alpha of the previous grid (normalised) on entry to a grid, alphaHat_m / alpha of this grid afterwards
    double cfin[NH];     // c of the last grid processed
#pragma unroll
    for (int h = 0; h < NH; h++) cfin[h] = 1;

    issue_pkg(0);
    // =============================================================== forward + read resampling
    for (int g = 0; g < T; g++) {
        wait_pkg(g);
        if (g + 1 < T) issue_pkg(g + 1);
        const int s = g & 1;
        const int r0 = rs[g], r1 = rs[g + 1];
        const int n_g = r1 - r0;
        const bool has = n_g > 0;
        double* eg = eGs + (size_t)(s * NH) * Kp;
        double c_old[NH];
#pragma unroll
        for (int h = 0; h < NH; h++) c_old[h] = ld_cg(J.c + h * T + g);
        double ab[NH][EPT];
        // beta of this grid goes straight into the ab registers (consumed after the forward step)
        if (has) {
#pragma unroll
            for (int h = 0; h < NH; h++) Col<NT, EPT>::load(ab[h], J.beta + ((size_t)h * T + g) * Kp, K, 0.0);
        }
        double cnew[NH];
        if (g == 0) {
            // rcpp_reinitialize_in_iterations
            double sv[NH];
#pragma unroll
            for (int h = 0; h < NH; h++) {
#pragma unroll
                for (int i = 0; i < EPT; i++) {
                    const int k = tid + i * NT;
                    am[h][i] = (k < K) ? one_over_K * eg[h * Kp + k] : 0.0;
                }
                sv[h] = Col<NT, EPT>::sum(am[h]);
            }
            bsum.run(sv);
#pragma unroll
            for (int h = 0; h < NH; h++) {
                cnew[h] = 1 / sv[h];
#pragma unroll
                for (int i = 0; i < EPT; i++) am[h][i] *= cnew[h];
            }
        } else {
            // rcpp_alpha_forward_one_QUILT_faster
            double sp[NH];
#pragma unroll
            for (int h = 0; h < NH; h++) sp[h] = Col<NT, EPT>::sum(am[h]);
            bsum.run(sp);
            const double x = J.tm[2 * (g - 1)], t1 = J.tm[2 * (g - 1) + 1];
            double sv[NH];
#pragma unroll
            for (int h = 0; h < NH; h++) {
                const double alphaConst = t1 * sp[h];
                const double jump = alphaConst * one_over_K;
#pragma unroll
                for (int i = 0; i < EPT; i++) {
                    const int k = tid + i * NT;
                    double v = 0.0;
                    if (k < K) {
                        v = x * am[h][i] + jump;
                        if (has) v = eg[h * Kp + k] * v;
                    }
                    am[h][i] = v;
                }
                sv[h] = Col<NT, EPT>::sum(am[h]);
            }
            bsum.run(sv);
#pragma unroll
            for (int h = 0; h < NH; h++) {
                double sc = 1 / (c_old[h] * sv[h]);
                cnew[h] = c_old[h] * sc;
                sc *= c_old[h];
#pragma unroll
                for (int i = 0; i < EPT; i++) am[h][i] *= sc;
            }
        }
        // am now holds alphaHat_t[:, g]; keep it as `ap` for the next grid unless reads change it
        bool changed = false;
        if (has) {
            const bool staged = (n_g <= SW_MAXR) && (J.ts[g + 1] - J.ts[g] <= SW_MAXTAB);
            const unsigned char* sm = smem + L.off_small[s];
            const ReadDesc* descp = staged ? reinterpret_cast<const ReadDesc*>(sm + L.small_desc) : J.desc + r0;
            const TabEnt* tabp = staged ? reinterpret_cast<const TabEnt*>(sm + L.small_tab) : J.tabs + J.ts[g];
            const double* Up = staged ? reinterpret_cast<const double*>(sm + L.small_U) : U + r0;
            const int32_t* Hp = staged ? reinterpret_cast<const int32_t*>(sm + L.small_H) : J.H + r0;
            const uint32_t tab_base = (uint32_t)J.ts[g];
            const uint32_t* wg = Wr + (g & 3) * Kp;  // this grid's allele words: most reads lie inside one 32-SNP word
            bool inited = false;
            P3 pC = {1, 1, 1};
            const P3 prior = {P.prior[0], P.prior[1], P.prior[2]};
            for (int ir = 0; ir < n_g; ir++) {
                // descriptor as two 16-byte words (no local-memory copy): off | cat,mode,nb,g0rel | b0,sel0..2 | ...
                const uint4 dq = *reinterpret_cast<const uint4*>(descp + ir);
                const int cat = dq.y & 0xff, mode = (dq.y >> 8) & 0xff, nb = (dq.y >> 16) & 0xff;
                if (NH == 2 && cat == 1) continue;  // diploid: uninformative reads are never visited (gibbs-nipt.cpp:815)
                const int g0rel = (int)(int8_t)(dq.y >> 24), b0 = dq.z & 0xff;
                const int r = r0 + ir;
                // which of the three regimes (gibbs-nipt.cpp:816-834)
                bool normal = true, init_mode = false, pass = false;
                if (iterative) {
                    if (iteration == 0) {
                        normal = false;
                        if (r < J.first_read)
                            pass = true;
                        else
                            init_mode = true;
                    } else if (iteration == 1 && r < J.first_read) {
                        normal = false;
                        init_mode = true;
                    }
                }
                if (!inited) {
                    // alphaHat_m = alpha, betaHat_m = beta, ab_m = alpha * beta ; pC = colsums
                    double sv[NH];
#pragma unroll
                    for (int h = 0; h < NH; h++) {
#pragma unroll
                        for (int i = 0; i < EPT; i++) ab[h][i] = am[h][i] * ab[h][i];
                        sv[h] = Col<NT, EPT>::sum(ab[h]);
                    }
                    bsum.run(sv);
                    pC.a = sv[0];
                    pC.b = sv[1];
                    if (NH == 3) pC.c = sv[NH - 1];
                    inited = true;
                }
                int hC = 0, hA1 = 1, hA2 = 2;
                P3 pA1 = pC, pA2 = pC;
                const TabEnt* tab = tabp + (dq.x - tab_base);
                const double* dcol = J.dense + (size_t)dq.x * Kp;
                // allele pattern of every haplotype over the read's SNPs (index into the read's emission table).
                // fast path: all SNPs inside this grid's word -> (word >> b0) & mask inline; otherwise the patterns
                // are materialised once into spat[] (each thread writes and later reads only its own elements)
                const bool fast = (mode == MODE_RUN) && g0rel == 0 && (b0 + nb <= 32);
                const uint32_t mask = (1u << nb) - 1u;
                if (!fast && !pass) {
                    if (mode == MODE_RUN) {
                        const uint32_t* wlo = Wr + ((g + g0rel) & 3) * Kp;
                        const uint32_t* whi = Wr + ((g + g0rel + 1) & 3) * Kp;
                        const bool cross = b0 + nb > 32;
#pragma unroll
                        for (int i = 0; i < EPT; i++) {
                            const int k = tid + i * NT;
                            if (k < K) {
                                const uint32_t lo = wlo[k];
                                const uint32_t hi = cross ? whi[k] : 0u;
                                spat[k] = (uint16_t)(__funnelshift_r(lo, hi, b0) & mask);
                            }
                        }
                    } else if (mode == MODE_GATHER) {
                        const uint8_t* sel = reinterpret_cast<const uint8_t*>(descp + ir) + 9;
#pragma unroll
                        for (int i = 0; i < EPT; i++) {
                            const int k = tid + i * NT;
                            if (k < K) {
                                uint32_t pt = 0;
                                for (int j = 0; j < nb; j++) {
                                    const int sj = sel[j];
                                    pt |= ((Wr[((g + (sj >> 5) - 1) & 3) * Kp + k] >> (sj & 31)) & 1u) << j;
                                }
                                spat[k] = (uint16_t)pt;
                            }
                        }
                    }
                }
#define QB_PAT(k) (fast ? ((wg[k] >> b0) & mask) : (uint32_t)spat[k])
                if (!pass) {
                    if (normal) {
                        hC = Hp[ir] - 1;
                        hA1 = (hC == 0) ? 1 : 0;
                        hA2 = (hC == 2) ? 1 : 2;
                    }
                    // K-long sums: normal  -> sum ab_C / e, sum ab_A1 * e, (sum ab_A2 * e)
                    //              init    -> sum ab_0 * e, sum ab_1 * e, (sum ab_2 * e)
                    double sv[NH];
#pragma unroll
                    for (int h = 0; h < NH; h++) sv[h] = 0;
#define QB_SUM_LOOP(XC, XA1, XA2, CMUL)                                                                  \
    _Pragma("unroll") for (int i = 0; i < EPT; i++) {                                                     \
        const int k = tid + i * NT;                                                                       \
        if (k < K) {                                                                                      \
            if (mode == MODE_DENSE) {                                                                     \
                const double E = dcol[k];                                                                 \
                sv[0] += CMUL ? XC[i] * E : XC[i] / E;                                                    \
                sv[1] += XA1[i] * E;                                                                      \
                if (NH == 3) sv[NH - 1] += XA2[i] * E;                                                    \
            } else {                                                                                      \
                const double2 te = *reinterpret_cast<const double2*>(tab + QB_PAT(k));                   \
                sv[0] = fma(XC[i], CMUL ? te.x : te.y, sv[0]);                                            \
                sv[1] = fma(XA1[i], te.x, sv[1]);                                                         \
                if (NH == 3) sv[NH - 1] = fma(XA2[i], te.x, sv[NH - 1]);                                  \
            }                                                                                             \
        }                                                                                                 \
    }
                    if (!normal) {
                        QB_SUM_LOOP(ab[0], ab[1], ab[NH - 1], true)
                    } else if (hC == 0) {
                        QB_SUM_LOOP(ab[0], ab[1], ab[NH - 1], false)
                    } else if (hC == 1) {
                        QB_SUM_LOOP(ab[1], ab[0], ab[NH - 1], false)
                    } else {
                        QB_SUM_LOOP(ab[NH - 1], ab[0], ab[1], false)
                    }
#undef QB_SUM_LOOP
                    bsum.run(sv);
                    if (normal) {
                        pA1.set(hC, sv[0]);
                        pA1.set(hA1, sv[1]);
                        if (NH == 3) pA2.set(hA2, sv[NH - 1]);
                        pA2.set(hC, sv[0]);
                    } else {
                        pC.a = sv[0];
                        pA1.b = sv[1];
                        if (NH == 3) pA2.c = sv[NH - 1];
                    }
                }
                const double prod_pC = pC.prod() * prior.get(hC);
                const double prod_pA1 = pA1.prod() * prior.get(hA1);
                const double prod_pA2 = pA2.prod() * prior.get(hA2);
                const double denom = prod_pC + prod_pA1 + prod_pA2;
                const double norm_pC = prod_pC / denom, norm_pA1 = prod_pA1 / denom, norm_pA2 = prod_pA2 / denom;
                const double chance = Up[ir];
                P3 cum = {0, 0, 0};
                cum.set(hC, norm_pC);
                cum.set(hA1, norm_pA1);
                cum.set(hA2, norm_pA2);
                const double x0 = cum.a, x1 = cum.b, x2 = cum.c;
                cum.b += cum.a;
                cum.c += cum.b;
                int hN = 0;
                if (chance < cum.c) hN = 2;
                if (chance < cum.b) hN = 1;
                if (chance < cum.a) hN = 0;
                if (((hN != hC) || init_mode) && !pass) {
                    changed = true;
                    if (tid == 0) J.H[r] = hN + 1;
                    // alphaHat_m / ab_m / eMatGrid of the old label are divided by the read's column, those of the new
                    // label multiplied (gibbs-nipt.cpp:1092-1110).  Divisions: q = a * (1/e) corrected by one fma
                    // residual step = the correctly rounded a / e (DESIGN.md "arithmetic"); dense columns divide.
#define QB_UPD_LOOP(HC, HN, DODIV)                                                                        \
    _Pragma("unroll") for (int i = 0; i < EPT; i++) {                                                     \
        const int k = tid + i * NT;                                                                       \
        if (k < K) {                                                                                      \
            double E, invE;                                                                               \
            if (mode == MODE_DENSE) {                                                                     \
                E = dcol[k];                                                                              \
                invE = 1 / E;                                                                             \
            } else {                                                                                      \
                const double2 te = *reinterpret_cast<const double2*>(tab + QB_PAT(k));                   \
                E = te.x;                                                                                 \
                invE = te.y;                                                                              \
            }                                                                                             \
            if (DODIV) {                                                                                  \
                am[HC][i] = div_by(am[HC][i], E, invE);                                                   \
                ab[HC][i] = div_by(ab[HC][i], E, invE);                                                   \
                if (HC < 2 || NH == 3) eg[HC * Kp + k] = div_by(eg[HC * Kp + k], E, invE);                \
            }                                                                                             \
            am[HN][i] *= E;                                                                               \
            ab[HN][i] *= E;                                                                               \
            if (HN < 2 || NH == 3) eg[HN * Kp + k] *= E;                                                  \
        }                                                                                                 \
    }
                    if (normal) {
                        if (hC == 0 && hN == 1) {
                            QB_UPD_LOOP(0, 1, true)
                        } else if (hC == 1 && hN == 0) {
                            QB_UPD_LOOP(1, 0, true)
                        } else if (NH == 3) {
                            if (hC == 0 && hN == 2) {
                                QB_UPD_LOOP(0, NH - 1, true)
                            } else if (hC == 1 && hN == 2) {
                                QB_UPD_LOOP(1, NH - 1, true)
                            } else if (hC == 2 && hN == 0) {
                                QB_UPD_LOOP(NH - 1, 0, true)
                            } else if (hC == 2 && hN == 1) {
                                QB_UPD_LOOP(NH - 1, 1, true)
                            }
                        }
                        const bool useA1 = (hN == hA1);
                        pC = useA1 ? pA1 : pA2;
                    } else {
                        if (hN == 0) {
                            QB_UPD_LOOP(0, 0, false)
                        } else if (hN == 1) {
                            QB_UPD_LOOP(0, 1, false)
                            pC = pA1;
                        } else if (NH == 3) {
                            QB_UPD_LOOP(0, NH - 1, false)
                            pC = pA2;
                        }
                    }
#undef QB_UPD_LOOP
#undef QB_PAT
                }
                if (record && tid == 0) {
                    J.xprob[3 * (size_t)r + 0] = x0;
                    J.xprob[3 * (size_t)r + 1] = x1;
                    J.xprob[3 * (size_t)r + 2] = x2;
                }
            }
            if (changed) {
                // gibbs-nipt.cpp:1262-1292: renormalise every haplotype's column and fold into c
                double sv[NH];
#pragma unroll
                for (int h = 0; h < NH; h++) sv[h] = Col<NT, EPT>::sum(am[h]);
                bsum.run(sv);
#pragma unroll
                for (int h = 0; h < NH; h++) {
                    const double alphaConst = 1 / sv[h];
                    cnew[h] *= alphaConst;
#pragma unroll
                    for (int i = 0; i < EPT; i++) am[h][i] *= alphaConst;
                }
            }
        }
#pragma unroll
        for (int h = 0; h < NH; h++) {
            Col<NT, EPT>::store(am[h], J.alpha + ((size_t)h * T + g) * Kp, K);
            if (changed) {
#pragma unroll
                for (int i = 0; i < EPT; i++) {
                    const int k = tid + i * NT;
                    if (k < K) st_stream(J.eG + ((size_t)h * T + g) * Kp + k, eg[h * Kp + k]);
                }
            }
            if (tid == 0) J.c[h * T + g] = cnew[h];
            cfin[h] = cnew[h];
        }
    }

    // =============================================================== backward (Rcpp_run_backward_haploid_QUILT_faster)
    // the eMatGrid columns changed above were written through the generic proxy; order them before the
    // bulk (async proxy) reads below
    __threadfence();
    fence_proxy_async_all();
    __syncthreads();
    {
        double b[NH][EPT];
#pragma unroll
        for (int h = 0; h < NH; h++) {
#pragma unroll
            for (int i = 0; i < EPT; i++) b[h][i] = (tid + i * NT < K) ? cfin[h] : 0.0;
            Col<NT, EPT>::store(b[h], J.beta + ((size_t)h * T + (T - 1)) * Kp, K);
        }
        // stage ring reused: step g needs eMatGrid[:, g + 1]
        auto issue_b = [&](int g) {
            const int s = g & 1;
            if (tid == 0) {
                const bool has1 = rs[g + 2] > rs[g + 1];
                if (has1) {
                    mbar_arrive_expect_tx(&bar[s], NH * Kp * 8);
#pragma unroll
                    for (int h = 0; h < NH; h++) bulk_g2s(eGs + (size_t)(s * NH + h) * Kp, J.eG + ((size_t)h * T + g + 1) * Kp, Kp * 8, &bar[s]);
                } else {
                    mbar_arrive(&bar[s]);
                }
            }
        };
        if (T >= 2) issue_b(T - 2);
        for (int g = T - 2; g >= 0; g--) {
            const int s = g & 1;
            if (s == 0) {
                mbar_wait(&bar[0], n_use0 & 1);
                n_use0++;
            } else {
                mbar_wait(&bar[1], n_use1 & 1);
                n_use1++;
            }
            __syncthreads();  // everyone is done with the other stage (step g + 1) before it is refilled
            if (g >= 1) issue_b(g - 1);
            const bool has1 = rs[g + 2] > rs[g + 1];
            const double* eg = eGs + (size_t)(s * NH) * Kp;
            const double t0 = J.tm[2 * g], t1 = J.tm[2 * g + 1];
            double cg[NH], sv[NH];
#pragma unroll
            for (int h = 0; h < NH; h++) {
                cg[h] = ld_cg(J.c + h * T + g);
                if (has1) {
#pragma unroll
                    for (int i = 0; i < EPT; i++) {
                        const int k = tid + i * NT;
                        if (k < K) b[h][i] = eg[h * Kp + k] * b[h][i];
                    }
                }
                sv[h] = Col<NT, EPT>::sum(b[h]);
            }
            bsum.run(sv);
#pragma unroll
            for (int h = 0; h < NH; h++) {
                const double x = t1 * sv[h] * one_over_K;
#pragma unroll
                for (int i = 0; i < EPT; i++) b[h][i] = (tid + i * NT < K) ? cg[h] * (x + t0 * b[h][i]) : 0.0;
                Col<NT, EPT>::store(b[h], J.beta + ((size_t)h * T + g) * Kp, K);
            }
        }
    }

    // =============================================================== epilogue: H_class, counts, -sum log c, underflow
    __syncthreads();
    for (int r = tid; r < R; r += NT) {
        const int h = J.H[r];
        atomicAdd(&cnt[h - 1], 1);
        if (record) {
            int hc;
            if (NH == 2 && J.desc[r].cat == 1) {
                hc = J.Hclass[r];
            } else {
                hc = classify_H(P, J.xprob[3 * (size_t)r], J.xprob[3 * (size_t)r + 1], J.xprob[3 * (size_t)r + 2]);
                J.Hclass[r] = hc;
            }
            atomicAdd(&cnt[3 + hc], 1);
        }
    }
    double sl[NH], sc[NH];
#pragma unroll
    for (int h = 0; h < NH; h++) {
        sl[h] = 0;
        sc[h] = 0;
        for (int g = tid; g < T; g += NT) {
            const double cv = ld_cg(J.c + h * T + g);
            sl[h] += log(cv);
            sc[h] += cv;
        }
    }
    bsum.run(sl);
    bsum.run(sc);
    __syncthreads();
    if (tid == 0) {
        double* lik = J.lik + (size_t)iteration * LIK_N;
        bool bad = false;
#pragma unroll
        for (int h = 0; h < NH; h++) {
            lik[h] = -sl[h];
            // gibbs-nipt.cpp:2959-2969: c1, c2 always; c3 only when ff == 0
            if (h < 2 || P.ff == 0) bad = bad || !isfinite(sc[h]);
        }
        for (int j = 0; j < 3; j++) lik[3 + j] = cnt[j];
        for (int j = 0; j < 8; j++) lik[6 + j] = cnt[3 + j];
        lik[14] = bad ? 1.0 : 0.0;
        if (bad) *J.underflow = 1;
    }
}

}  // namespace qb
