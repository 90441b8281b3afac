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
// registers); alpha of the previous grid and the working copies alphaHat_m / ab_m of the current grid live in
// registers for the whole grid.  eMatGrid and beta columns (two ring buffers with three rotating tenants) and the
// 32-SNP allele words of grids g-1 .. g+2 are staged in shared memory by 1-D bulk async copies (TMA engine,
// mbarrier completion); the small per-grid read metadata (descriptors, emission tables, labels, uniforms, per-grid
// scalars) by cp.async.  alpha / beta / changed eMatGrid columns leave with streaming stores.  Per-read K-long sums
// use a shuffle butterfly + one shared-memory exchange; every thread ends with bit-identical totals and takes the
// label decision redundantly.  Shared-memory columns have stride KA = NT * EPT >= K so the per-read loops need no
// bounds checks (padding elements carry alpha = ab = 0).  For K > 4096 a job runs on a two-CTA cluster (CL = 2).
#pragma once

#include "device_common.cuh"
#include "types.h"

// -DQB_CLK=1 builds an instrumented variant (tools/build_variant.py): thread 0 of every CTA accumulates clock64()
// deltas per phase of the sweep and CTA 0 prints them after sweep QB_CLK_IT.  Never defined in the product build.
#ifndef QB_CLK
#define QB_CLK 0
#endif
#ifndef QB_CLK_IT
#define QB_CLK_IT 5
#endif
#ifndef QB_CLK_TID
#define QB_CLK_TID 0  // the thread whose clock is read (its warp's view of every phase)
#endif
#if QB_CLK
#define QB_T(i)                                 \
    do {                                        \
        if (tid == QB_CLK_TID) {                \
            const long long n_ = clock64();     \
            clk[i] += n_ - tlast;               \
            tlast = n_;                         \
        }                                       \
    } while (0)
#define QB_N(i)                           \
    do {                                  \
        if (tid == QB_CLK_TID) clk[i]++;  \
    } while (0)
#else
#define QB_T(i) ((void)0)
#define QB_N(i) ((void)0)
#endif

namespace qb {

constexpr int SW_MAXR = 48;     // reads staged in shared memory at a time (longer grids are processed in chunks)
constexpr int SW_MAXTAB = 704;  // table entries staged at a time (>= 2^NBMAX, so any table-mode read fits); sized so that
                                // ~97 % of the all-SNP grids of the benchmark (311 +- 205 entries) are staged with their package
constexpr int SW_VMAX = 4;
constexpr int SW_BMAX = 16;  // reads decided per round of the batched resampler
static_assert(SW_MAXTAB >= (1 << NBMAX), "a single emission table must fit the staging buffer");

struct SweepSmemLayout {
    int off_bar, off_red, off_cnt, off_sc, off_small[2], off_pat, off_part, off_rec, off_cls, off_W, off_eG, total;
    int small_desc, small_tab, small_U, small_H;
};
__host__ __device__ inline SweepSmemLayout sweep_smem_layout(int KA, int NH, int NT, int CL = 1) {
    SweepSmemLayout L;
    int o = 0;
    L.off_bar = o;
    o += 64;
    L.off_red = o;
    o += 2 * CL * SW_VMAX * (NT / 32) * 8;
    o = (o + 127) & ~127;
    L.off_cnt = o;
    o += 128;
    L.off_sc = o;  // per-stage scalar packages: 2 forward + 3 backward, 64 bytes each
    o += 5 * 64 + 64;
    L.small_desc = 0;
    L.small_tab = SW_MAXR * 32;
    L.small_U = L.small_tab + SW_MAXTAB * 16;
    L.small_H = L.small_U + SW_MAXR * 8;
    const int small_total = (L.small_H + SW_MAXR * 4 + 127) & ~127;
    L.off_small[0] = o;
    o += small_total;
    L.off_small[1] = o;
    o += small_total;
    L.off_pat = o;  // (gather-mode pattern scratch: the host turns such reads into dense columns, nothing to reserve)
    L.off_part = o;
    L.off_rec = o;  // chunking scratch of grids whose reads exceed one staging buffer
    o += 64;
    L.off_cls = o;  // class path (diploid, one CTA per job): class totals / column factors [2][CLS_AS] + per-warp scan carries [NT / 32][4]
    if (NH == 2 && CL == 1) o += 2 * CLS_AS * 8 + (NT / 32) * 32;
    o = (o + 127) & ~127;
    L.off_W = o;
    o += 4 * KA * 4;
    L.off_eG = o;
    o += 2 * NH * KA * 8;
    L.total = o;
    return L;
}

template <int NT, int CL = 1>
struct BlockSumV {
    static constexpr int NW = NT / 32;  // power of two (4, 8, 16)
    double* scratch;                    // [2][CL][SW_VMAX][NW]
    int phase;
    uint32_t crank;
    __device__ __forceinline__ BlockSumV(double* s, uint32_t rank = 0) : scratch(s), phase(0), crank(rank) {}
    // CL == 2: the K states of a job are split over the two CTAs of a cluster.  Every warp pushes its partial into the
    // scratch of BOTH CTAs (own shared memory + distributed shared memory of the peer), one cluster barrier replaces
    // __syncthreads, and every thread of both CTAs adds the 2 * NW partials in the same fixed order (rank 0 first),
    // so the two CTAs hold bit-identical totals and take identical decisions.
    template <int V>
    __device__ __forceinline__ void run_cluster(double (&v)[V]) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int i = 0; i < V; i++) {
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], d);
        }
        double* buf = scratch + phase * (CL * SW_VMAX * NW);
        phase ^= 1;
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < V; i++) {
                double* slot = buf + (crank * SW_VMAX + i) * NW + warp;
                *slot = v[i];
                st_dsmem(slot, crank ^ 1u, v[i]);
            }
        }
        cluster_sync_all();
#pragma unroll
        for (int i = 0; i < V; i++) {
            double p[CL * NW];
#pragma unroll
            for (int r = 0; r < CL; r++)
#pragma unroll
                for (int w = 0; w < NW; w++) p[r * NW + w] = buf[(r * SW_VMAX + i) * NW + w];
#pragma unroll
            for (int st = 1; st < CL * NW; st <<= 1) {
#pragma unroll
                for (int w = 0; w + st < CL * NW; w += 2 * st) p[w] += p[w + st];
            }
            v[i] = p[0];
        }
    }
    // Every thread returns the same, bit-identical totals: xor butterfly inside each warp, one shared-memory slot per
    // warp, then every warp butterflies the NW partials again (same operation order in all warps).  Two alternating
    // scratch buffers make one __syncthreads per call enough.
    template <int V>
    __device__ __forceinline__ void run(double (&v)[V]) {
        if (CL > 1) {
            run_cluster(v);
            return;
        }
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (V == 2 && NW > 1) {
            // two values: the lower half-warp reduces v[0], the upper half v[1] (half the shuffles of two butterflies)
            const bool upper = (lane & 16) != 0;
            const double send = upper ? v[0] : v[1];
            double x = (upper ? v[1] : v[0]) + __shfl_xor_sync(0xffffffffu, send, 16);
#pragma unroll
            for (int d = 8; d >= 1; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
            double* buf = scratch + phase * (SW_VMAX * NW);
            phase ^= 1;
            if ((lane & 15) == 0) buf[(lane >> 4) * NW + warp] = x;
            __syncthreads();
            // every thread reads all warp partials (broadcast 16-byte loads) and adds them in one fixed tree: ~3 dependent
            // adds instead of log2(NW) + 1 shuffle round trips
            double p0[NW], p1[NW];
#pragma unroll
            for (int w = 0; w < NW; w += 2) {
                const double2 a = *reinterpret_cast<const double2*>(buf + w);
                const double2 b = *reinterpret_cast<const double2*>(buf + NW + w);
                p0[w] = a.x;
                p0[w + 1] = a.y;
                p1[w] = b.x;
                p1[w + 1] = b.y;
            }
#pragma unroll
            for (int st = 1; st < NW; st <<= 1) {
#pragma unroll
                for (int w = 0; w + st < NW; w += 2 * st) {
                    p0[w] += p0[w + st];
                    p1[w] += p1[w + st];
                }
            }
            v[0] = p0[0];
            v[1] = p1[0];
            return;
        }
#pragma unroll
        for (int i = 0; i < V; i++) {
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], d);
        }
        if (NW == 1) return;
        double* buf = scratch + phase * (SW_VMAX * NW);
        phase ^= 1;
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


enum ReadKind : int { KIND_NORMAL = 0, KIND_INIT = 1, KIND_PASS = 2 };

struct Decision {
    int hN;
    bool change;
    P3 x;      // label probabilities placed by label (H_class is derived from them)
    P3 pCnew;  // column sums of ab_m after the decision
};

// the scalar part of sample_reads_in_grid for one read (gibbs-nipt.cpp:998-1046, :1112-1134):
// sv = {sum ab_C / e, sum ab_A1 * e, sum ab_A2 * e} (normal) or {sum ab_0 * e, sum ab_1 * e, sum ab_2 * e} (initialisation)
// a / b from rb = RN(1 / b): same correction step as div_by, i.e. the correctly rounded quotient
__device__ __forceinline__ double quot(double a, double b, double rb) {
    const double q = a * rb;
    return fma(fma(-q, b, a), rb, q);
}

// Diploid, normal regime, decisive case: the label follows from comparing chance * denom with the unnormalised
// probability of label 0; a guard band (1e-11 relative, five orders above the rounding of the exact path) and the
// chance ~ 1 corner send everything else to the exact code, so the labels are those of decide_read bit for bit.
// The normalised probabilities (only needed for H_class) are formed in the sweep epilogue from the recorded products.
struct FastDecision {
    bool decided;
    int hN;
    double prod_pC, prod_pA1, prod_pA2;
};
__device__ __forceinline__ FastDecision decide_diploid_fast(const P3& pC_in, double sv0, double sv1, int hC_in, double chance, const P3& prior) {
    FastDecision F;
    const bool c0 = hC_in == 0;
    const double pa = c0 ? pC_in.b : pC_in.a;
    F.prod_pC = ((pC_in.a * pC_in.b) * pC_in.c) * (c0 ? prior.a : prior.b);
    F.prod_pA1 = ((c0 ? sv0 * sv1 : sv1 * sv0) * pC_in.c) * (c0 ? prior.b : prior.a);
    F.prod_pA2 = ((c0 ? sv0 * pa : pa * sv0) * pC_in.c) * prior.c;
    const double denom = F.prod_pC + F.prod_pA1 + F.prod_pA2;
    const double P0 = c0 ? F.prod_pC : F.prod_pA1;
    const double t = chance * denom;
    const double margin = 1e-11 * denom;
    F.decided = false;
    F.hN = 0;
    if (t < P0 - margin) {
        F.decided = true;
        F.hN = 0;
    } else if (t > P0 + margin && chance < 1 - 1e-9 && F.prod_pA2 == 0) {
        F.decided = true;
        F.hN = 1;
    }
    return F;
}

template <int NH>
__device__ __forceinline__ Decision decide_read(const P3& pC_in, const double (&sv)[NH], int hC_in, int kind, double chance, const P3& prior) {
    Decision D;
    if (NH == 2 && kind == KIND_NORMAL) {
        // diploid fast path, same values as the general code below: label 2 has prior 0 and pC.c stays 1
        const bool c0 = hC_in == 0;
        const double pc = c0 ? pC_in.a : pC_in.b, pa = c0 ? pC_in.b : pC_in.a;  // current / alternative column sums
        const double prod_pC = ((pC_in.a * pC_in.b) * pC_in.c) * (c0 ? prior.a : prior.b);
        const double prod_pA1 = ((c0 ? sv[0] * sv[1] : sv[1] * sv[0]) * pC_in.c) * (c0 ? prior.b : prior.a);
        const double prod_pA2 = ((c0 ? sv[0] * pa : pa * sv[0]) * pC_in.c) * prior.c;
        (void)pc;
        const double denom = prod_pC + prod_pA1 + prod_pA2;
        const double rd = 1 / denom;
        const double norm_pC = quot(prod_pC, denom, rd), norm_pA1 = quot(prod_pA1, denom, rd), norm_pA2 = quot(prod_pA2, denom, rd);
        D.x.a = c0 ? norm_pC : norm_pA1;
        D.x.b = c0 ? norm_pA1 : norm_pC;
        D.x.c = norm_pA2;
        const double cb = D.x.b + D.x.a, cc = D.x.c + cb;
        int hN = 0;
        if (chance < cc) hN = 2;
        if (chance < cb) hN = 1;
        if (chance < D.x.a) hN = 0;
        D.hN = hN;
        D.change = hN != hC_in;
        D.pCnew = pC_in;
        if (D.change) {
            if (hN == (c0 ? 1 : 0)) {
                D.pCnew.a = c0 ? sv[0] : sv[1];
                D.pCnew.b = c0 ? sv[1] : sv[0];
            } else {
                D.pCnew.a = c0 ? sv[0] : pC_in.a;
                D.pCnew.b = c0 ? pC_in.b : sv[0];
            }
        }
        return D;
    }
    int hC = 0, hA1 = 1, hA2 = 2;
    P3 pC = pC_in, pA1 = pC_in, pA2 = pC_in;
    if (kind == KIND_NORMAL) {
        hC = hC_in;
        hA1 = (hC == 0) ? 1 : 0;
        hA2 = (hC == 2) ? 1 : 2;
        pA1.set(hC, sv[0]);
        pA1.set(hA1, sv[1]);
        if (NH == 3) pA2.set(hA2, sv[NH - 1]);
        pA2.set(hC, sv[0]);
    } else if (kind == KIND_INIT) {
        pC.a = sv[0];
        pA1.b = sv[1];
        if (NH == 3) pA2.c = sv[NH - 1];
    }
    const double prod_pC = pC.prod() * prior.get(hC);
    const double prod_pA1 = pA1.prod() * prior.get(hA1);
    const double prod_pA2 = pA2.prod() * prior.get(hA2);
    const double denom = prod_pC + prod_pA1 + prod_pA2;
    double norm_pC, norm_pA1, norm_pA2;
    if (denom > 0 && denom < 1e300) {
        // one reciprocal + three corrected products = the correctly rounded quotients (see quot), a third of the division work
        const double rd = 1 / denom;
        norm_pC = quot(prod_pC, denom, rd);
        norm_pA1 = quot(prod_pA1, denom, rd);
        norm_pA2 = quot(prod_pA2, denom, rd);
    } else {  // zero / non-finite sums (underflow): plain IEEE divisions, whatever they give
        norm_pC = prod_pC / denom;
        norm_pA1 = prod_pA1 / denom;
        norm_pA2 = prod_pA2 / denom;
    }
    P3 cum = {0, 0, 0};
    cum.set(hC, norm_pC);
    cum.set(hA1, norm_pA1);
    cum.set(hA2, norm_pA2);
    D.x = cum;
    cum.b += cum.a;
    cum.c += cum.b;
    int hN = 0;
    if (chance < cum.c) hN = 2;
    if (chance < cum.b) hN = 1;
    if (chance < cum.a) hN = 0;
    D.hN = hN;
    D.change = ((hN != hC) || kind == KIND_INIT) && kind != KIND_PASS;
    D.pCnew = pC;
    if (D.change) {
        if (kind == KIND_NORMAL)
            D.pCnew = (hN == hA1) ? pA1 : pA2;
        else if (hN == 1)
            D.pCnew = pA1;
        else if (hN == 2)
            D.pCnew = pA2;
    }
    return D;
}

// N values per lane -> after the call v[0] of lane L holds the 32-lane total of value index
// wtr_index<N>(L); 4 (N = 8) or 2 (N = 16) lanes hold each value.  31/N as many shuffles as N butterflies.
template <int N>
__device__ __forceinline__ void warp_transpose_reduce(double (&v)[N], int lane) {
    int n = N;
#pragma unroll
    for (int d = 16; n > 1; d >>= 1, n >>= 1) {
        const bool upper = (lane & d) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; i++) {
            const double send = upper ? v[i] : v[i + n / 2];
            const double keep = upper ? v[i + n / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, d);
        }
    }
#pragma unroll
    for (int d = 16 / N; d >= 1; d >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], d);
}
template <int N>
__device__ __forceinline__ int wtr_index(int lane) {
    return N == 8 ? (lane >> 2) : (lane >> 1);
}

// where a read's emission value comes from, per haplotype k (uniform over the CTA for a given read)
struct ESrc {
    const uint32_t* wlo;   // RUN: word holding the first SNP
    const uint32_t* whi;   // RUN: next word (only read when the run crosses)
    const uint16_t* spat;  // GATHER: materialised patterns
    const TabEnt* tab;     // table (shared memory)
    const TabEnt* dcol;    // DENSE: K-long {E, 1/E} column (global memory)
    uint32_t b0, mask;
    bool cross;
};
struct EV {
    double E, invE;
};
template <int SRC>  // 0: run inside one word, 1: run crossing into the next word, 2: materialised patterns, 3: dense column
__device__ __forceinline__ EV emission_at(const ESrc& S, int k) {
    EV r;
    if (SRC == 3) {
        double2 te;
        asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(te.x), "=d"(te.y) : "l"(S.dcol + k));
        r.E = te.x;
        r.invE = te.y;
        return r;
    }
    uint32_t pat;
    if (SRC == 0)
        pat = (S.wlo[k] >> S.b0) & S.mask;
    else if (SRC == 1)
        pat = __funnelshift_r(S.wlo[k], S.whi[k], S.b0) & S.mask;
    else
        pat = S.spat[k];
    const double2 te = *reinterpret_cast<const double2*>(S.tab + pat);
    r.E = te.x;
    r.invE = te.y;
    return r;
}

// Column update after a label change (gibbs-nipt.cpp:1092-1110): divide the old label's alphaHat_m / ab_m / eMatGrid
// by the read's column, multiply the new label's.  Elements are handled in chunks of UPD_CH with all shared-memory
// loads (allele word, table entry, eMatGrid values) issued before any store, so the loads of a chunk pipeline instead
// of serialising behind the previous element's store (the compiler cannot prove table and eMatGrid do not alias).
constexpr int UPD_CH = 4;
template <int NT, int EPT, int NH, int SRC, int HC, int HN, bool DODIV>
__device__ __forceinline__ void upd_loop(double (&am)[NH][EPT], double (&ab)[NH][EPT], double* eg, int KA, const ESrc& S, int K, int tid) {
    constexpr bool EGC = DODIV && (HC < 2 || NH == 3);
    constexpr bool EGN = (HN < 2 || NH == 3);
#pragma unroll
    for (int i0 = 0; i0 < EPT; i0 += UPD_CH) {
        EV ev[UPD_CH];
        double gc[UPD_CH], gn[UPD_CH];
#pragma unroll
        for (int j = 0; j < UPD_CH; j++) {
            const int k = tid + (i0 + j) * NT;
            if (SRC != 3 || k < K) {
                ev[j] = emission_at<SRC>(S, k);
                if (EGC) gc[j] = eg[HC * KA + k];
                if (EGN) gn[j] = eg[HN * KA + k];
            } else {
                ev[j].E = 1.0;
                ev[j].invE = 1.0;
                gc[j] = 0.0;
                gn[j] = 0.0;
            }
        }
#pragma unroll
        for (int j = 0; j < UPD_CH; j++) {
            const int i = i0 + j;
            if (DODIV) {
                am[HC][i] = div_by(am[HC][i], ev[j].E, ev[j].invE);
                ab[HC][i] = div_by(ab[HC][i], ev[j].E, ev[j].invE);
                if (EGC) gc[j] = div_by(gc[j], ev[j].E, ev[j].invE);
            }
            am[HN][i] *= ev[j].E;
            ab[HN][i] *= ev[j].E;
            if (EGN) gn[j] *= ev[j].E;
        }
#pragma unroll
        for (int j = 0; j < UPD_CH; j++) {
            const int k = tid + (i0 + j) * NT;
            if (SRC != 3 || k < K) {
                if (EGC) eg[HC * KA + k] = gc[j];
                if (EGN) eg[HN * KA + k] = gn[j];
            }
        }
    }
}

// CL = 2: the job's K states are split over the two CTAs of a thread-block cluster (K > NT * EPT): CTA `crank` owns
// k in [crank * KA, crank * KA + KA); block sums meet through distributed shared memory (BlockSumV<NT, 2>), both CTAs
// take every decision redundantly.
template <int NT, int EPT, int NH, int CL = 1, bool CLSP = false>
__global__ void __launch_bounds__(NT, (NT * EPT <= 512 ? 4 : (NT * EPT <= 1024 ? 2 : 1))) k_sweep(BatchParams P, const JobDev* __restrict__ jobs, int iteration, int store_alpha) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ JobDev Js;
    constexpr int KA = NT * EPT;
    const int tid = threadIdx.x;
    if (tid == 0) Js = jobs[blockIdx.x / CL];
    __syncthreads();
    const JobDev& J = Js;
    if (*J.underflow) return;  // (both CTAs of a cluster read the same flag)
    const uint32_t crank = (CL > 1) ? cluster_ctarank() : 0u;
    const int kbase = (CL > 1) ? (int)crank * KA : 0;                     // first haplotype state of this CTA
    const int K = (CL > 1) ? max(0, min(P.K - kbase, KA)) : P.K;          // states of this CTA (all bounds checks below are local)
    const int Kp = P.Kp;                                                  // column stride in HBM
    const int Kpl = (CL > 1) ? max(0, min(P.Kp - kbase, KA)) : P.Kp;      // padded states of this CTA (bulk-copy length)
    // (per-job scalars — labels, c, label-probability records, the sweep record — are written by both CTAs: the values
    //  are bit-identical, and each CTA only ever reads back what it wrote itself)
    const int T = P.T, R = J.R;
    // class path: reads of a grid whose haplotypes fall into <= CLS_MAX classes are decided on class totals (classes.cuh)
    // (CLSP: the instance for calls with many reads per grid — common-SNP calls; calls with few reads per grid run the
    //  instance without it, whose K-long path keeps the registers to itself)
    constexpr bool CLS_ON = CLSP && (NH == 2 && CL == 1);
    const SweepSmemLayout L = sweep_smem_layout(KA, NH, NT, CL);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.off_bar);
    BlockSumV<NT, CL> bsum(reinterpret_cast<double*>(smem + L.off_red), crank);
    int* cnt = reinterpret_cast<int*>(smem + L.off_cnt);
    uint32_t* Wr = reinterpret_cast<uint32_t*>(smem + L.off_W);          // [4][KA] ring of allele words
    uint16_t* spat = reinterpret_cast<uint16_t*>(smem + L.off_pat);      // [KA] allele patterns of gather-mode reads
    double* eGs = reinterpret_cast<double*>(smem + L.off_eG);            // [2][NH][KA]
    const bool iterative = (P.flags & QUILT_F_GIBBS_INITIALIZE_ITERATIVELY) != 0;
    const bool record = (P.flags & QUILT_F_RECORD_READ_SET) != 0;
    const double one_over_K = P.one_over_K;
    const int32_t* __restrict__ rs = J.rs;
    const int32_t* __restrict__ tsG = J.ts;
    const double* __restrict__ tmG = J.tm;
    const double* __restrict__ U = J.runif_reads + (size_t)iteration * R;
    double* __restrict__ alphaG = J.alpha + kbase;
    double* __restrict__ betaG = J.beta + kbase;
    double* __restrict__ eGg = J.eG + kbase;
    double* cG = J.c;

    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_init(&bar[2], 1);
        mbar_init(&bar[3], 1);  // eMatGrid columns into ring buffer 0 / 1 (forward)
        mbar_init(&bar[4], 1);
        mbar_init(&bar[5], 1);  // beta columns (forward)
        fence_barrier_init();
    }
    if (tid < 16) cnt[tid] = 0;
#if QB_CLK
    __shared__ long long clk[24];
    if (tid < 24) clk[tid] = 0;
    long long tlast = clock64();
#endif
    __syncthreads();
    const int n_tab_total = tsG[T];
    uint32_t n_use0 = 0, n_use1 = 0, n_use2 = 0;  // completed uses of each stage barrier -> wait parity
    uint32_t n_useE0 = 0, n_useE1 = 0, n_useB = 0;

    // staging of reads [ra, ra + n) with table entries [ta, ta + nt) into stage buffer s
    auto stage_small = [&](int s, int ra, int n, int ta, int nt, bool async) {
        // warp 0 issues the bulk copies of the package; the other warps share the cp.async staging so that no warp
        // reaches the next block barrier much later than the rest
        unsigned char* sm = smem + L.off_small[s];
        constexpr int NS = (NT > 32) ? NT - 32 : NT;
        const int t2 = (NT > 32) ? tid - 32 : tid;
        if (t2 >= 0) {
            const unsigned char* gd = reinterpret_cast<const unsigned char*>(J.desc + ra);
            for (int i = t2; i < n * 2; i += NS) cp_async16(sm + L.small_desc + i * 16, gd + i * 16);
            const unsigned char* gt = reinterpret_cast<const unsigned char*>(J.tabs + ta);
            for (int i = t2; i < nt; i += NS) cp_async16(sm + L.small_tab + i * 16, gt + i * 16);
            for (int i = t2; i < n; i += NS) {
                cp_async8(sm + L.small_U + i * 8, U + ra + i);
                cp_async4(sm + L.small_H + i * 4, J.H + ra + i);
            }
        }
        cp_async_commit();
        if (!async) {
            cp_async_wait_all();
            __syncthreads();
        }
    };
    // ---- one package = everything grid g needs: eMatGrid columns, the allele words of grid g+1 (ring of 4:
    //      grid g uses words g-1 .. g+1 while package g+1 is already filling word g+2), and the read
    //      metadata of the grid when it fits one staging buffer
    auto issue_pkg = [&](int g, int r0, int r1, int t0, int fcn, int fct) {
        const int s = g & 1;
        const int n_g = r1 - r0;
        {
            // scalars the serial chain needs at grid g: first reads / table offsets of the NEXT grid (so that its
            // package can be issued without a dependent global load), the transition pair into g, the previous c of g
            // one predicated 4-byte cp.async per lane of warp 1 (no divergent single-thread blocks: warp 0 already
            // carries the bulk-copy issue, and everything a single warp does alone delays the next block barrier)
            // block layout (96 bytes): ginfo rows g + 1 and g + 2 {first read, table offset, reads / table end of the
            // first staging chunk}, transition pair, previous c
            unsigned char* sc = smem + L.off_sc + s * 96;
            const int q = (NT > 32) ? tid - 32 : tid;
            if (q >= 0 && q < 12 + 2 * NH + (CLS_ON ? 1 : 0)) {
                const int32_t* src;
                bool ok = true;
                if (CLS_ON && q == 12 + 2 * NH) {
                    src = J.cinfo + g + 1;  // classes of the next grid
                } else if (q < 8) {
                    src = J.ginfo + 4 * (g + 1) + q;
                    ok = g + 1 + (q >> 2) <= T;
                } else if (q < 12) {
                    src = reinterpret_cast<const int32_t*>(tmG + 2 * (g - 1)) + (q - 8);
                    ok = g >= 1;
                } else {
                    const int hh = (q - 12) >> 1;
                    src = reinterpret_cast<const int32_t*>(cG + hh * T + g) + ((q - 12) & 1);
                }
                if (ok) cp_async4(sc + 4 * q, src);
            }
        }
        if (tid == 0) {
            uint32_t bytes = 0;
            if (g == 0) bytes += Kpl * 4;
            if (g + 1 < T) bytes += Kpl * 4;
            if (bytes > 0)
                mbar_arrive_expect_tx(&bar[s], bytes);
            else
                mbar_arrive(&bar[s]);
            if (g == 0) bulk_g2s(Wr, J.W + kbase, Kpl * 4, &bar[s]);
            if (g + 1 < T) bulk_g2s(Wr + ((g + 1) & 3) * KA, J.W + kbase + (size_t)(g + 1) * Kp, Kpl * 4, &bar[s]);
        }
        // the first staging chunk of the grid (all of its reads unless the grid is over-full; host-computed packing)
        if (n_g > 0)
            stage_small(s, r0, fcn, t0, fct - t0, true);
        else
            cp_async_commit();
        // Pull the FOLLOWING grid's read metadata towards L2 (a window from its first read / table entry on), so that
        // the cp.async package issued for it one grid later is served from L2 (measured: ~1 % of the sweep).
        if (!(P.dbg & 8)) {
            constexpr int PF_READS = 32, PF_TAB = 512;
            const int t1 = fct;  // (tables of the following grid start at or after the end of this grid's first chunk)
            const int nd = min(PF_READS, R - r1), ntb = min(PF_TAB, n_tab_total - t1);
            // one bulk L2 prefetch per array, each issued by lane 0 of a different warp (warps 0 and 1 carry the bulk-copy issue and
            // the scalar package)
            const int wq = (NT >= 192) ? tid - 64 : tid;
            if (wq == 0) {
                if (nd > 0) bulk_prefetch_l2_range(J.desc + r1, (size_t)nd * 32);
            } else if (wq == 32) {
                if (ntb > 0) bulk_prefetch_l2_range(J.tabs + t1, (size_t)ntb * 16);
            } else if (wq == 64) {
                if (nd > 0) bulk_prefetch_l2_range(U + r1, (size_t)nd * 8);
            } else if (wq == 96) {
                if (nd > 0) bulk_prefetch_l2_range(J.H + r1, (size_t)nd * 4);
            }
        }
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
    // The two big ring buffers alternate between three tenants (all by bulk copy, i.e. off the LSU and not ordered by
    // the block barriers): buffer g & 1 holds eMatGrid[:, g] for the whole of grid g; the other buffer holds
    // beta[:, g] from the top of grid g until alpha * beta has been formed, and then receives eMatGrid[:, g + 1].
    auto issue_eG = [&](int g) {  // caller: every thread is done with buffer g & 1
        if (tid == 0) {
            const int s = g & 1;
            bulk_wait_read_all();  // (a changed eMatGrid column may still be leaving this buffer by bulk store)
            fence_proxy_async();
            mbar_arrive_expect_tx(&bar[3 + s], NH * Kpl * 8);
#pragma unroll
            for (int h = 0; h < NH; h++) bulk_g2s(eGs + (size_t)(s * NH + h) * KA, eGg + ((size_t)h * T + g) * Kp, Kpl * 8, &bar[3 + s]);
        }
    };
    auto wait_eG = [&](int g) {
        if (g & 1) {
            mbar_wait(&bar[4], n_useE1 & 1);
            n_useE1++;
        } else {
            mbar_wait(&bar[3], n_useE0 & 1);
            n_useE0++;
        }
    };
    auto issue_beta = [&](int g) {  // into the buffer eMatGrid[:, g - 1] has left
        if (tid == 0) {
            const int s = (g & 1) ^ 1;
            bulk_wait_read_all();  // (eMatGrid[:, g - 1], if it changed, leaves that buffer by bulk store)
            fence_proxy_async();
            mbar_arrive_expect_tx(&bar[5], NH * Kpl * 8);
#pragma unroll
            for (int h = 0; h < NH; h++) bulk_g2s(eGs + (size_t)(s * NH + h) * KA, betaG + ((size_t)h * T + g) * Kp, Kpl * 8, &bar[5]);
        }
    };


    double am[NH][EPT];  // alpha of the previous grid (normalised) on entry to a grid, alphaHat_m / alpha of this grid afterwards
    double cfin[NH];     // c of the last grid processed
#pragma unroll
    for (int h = 0; h < NH; h++) cfin[h] = 1;

    // per-grid scalars arrive with the package (cp.async into the stage's scalar block): nothing on the serial chain
    // waits for a dependent global load
    int cur_r0 = J.ginfo[0], cur_t0 = J.ginfo[1], cur_fcn = J.ginfo[2], cur_fct = J.ginfo[3];
    int cur_nC = (CLS_ON && !(P.dbg & 16)) ? J.cinfo[0] : 0;
    issue_pkg(0, cur_r0, rs[1], cur_t0, cur_fcn, cur_fct);
    issue_eG(0);
    // =============================================================== forward + read resampling
    for (int g = 0; g < T; g++) {
        QB_T(11);
        wait_pkg(g);
        QB_T(0);
        const int s = g & 1;
        const int32_t* sci = reinterpret_cast<const int32_t*>(smem + L.off_sc + s * 96);
        const double* scd = reinterpret_cast<const double*>(sci);
        const int r0 = cur_r0, r1 = sci[0];
        const int ts0 = cur_t0, ts1 = sci[1];
        const int fcn = cur_fcn, fct = cur_fct;        // first staging chunk of this grid (already in shared memory)
        const int nx_fcn = sci[2], nx_fct = sci[3];    // ... of the next grid
        const int nx_r1 = sci[4];
        const int nC = cur_nC;
        if (CLS_ON) cur_nC = (P.dbg & 16) ? 0 : sci[12 + 2 * NH];
        const double tm_x = scd[4], tm_t1 = scd[5];
        const int n_g = r1 - r0;
        const bool has = n_g > 0;
        double* eg = eGs + (size_t)(s * NH) * KA;
        double c_old[NH];
#pragma unroll
        for (int h = 0; h < NH; h++) c_old[h] = scd[6 + h];
        cur_r0 = r1;
        cur_t0 = ts1;
        cur_fcn = nx_fcn;
        cur_fct = nx_fct;
        double ab[NH][EPT];
        const bool next_has = (g + 1 < T) && (nx_r1 > r1);
        if (has || g == 0) wait_eG(g);
        // beta of this grid arrives by bulk copy in the other ring buffer (consumed after the forward step)
        bool beta_pending = false, eG_next_issued = !next_has;
        if (has) {
            issue_beta(g);
            beta_pending = true;
        } else if (next_has) {
            issue_eG(g + 1);
            eG_next_issued = true;
        }
        if (g + 1 < T) {
            if (!(P.dbg & 2)) issue_pkg(g + 1, r1, nx_r1, ts1, nx_fcn, nx_fct);
            if (P.dbg & 4) cp_async_wait_all();  // experiment: does the next barrier already wait for the cp.async package?
            // pull the next grid's beta columns towards L2: one bulk prefetch per haplotype, lane 0 of warp h + 1
            if (nx_r1 > r1 && !(P.dbg & 1)) {
#pragma unroll
                for (int h = 0; h < NH; h++)
                    if (tid == 32 * ((h + 1) % (NT / 32))) bulk_prefetch_l2(betaG + ((size_t)h * T + g + 1) * Kp, (uint32_t)Kpl * 8);
            }
        }
        // class path of this grid (classes.cuh)?  Its layouts are pulled towards L2 one grid ahead and loaded after the forward step.
        bool use_cls = false;
        if (CLS_ON) {
            use_cls = has && nC > 0 && !(iterative && (iteration == 0 || (iteration == 1 && r0 < J.first_read)));
            if (cur_nC > 0 && g + 1 < T) {
                // sorted list (2 KA bytes), thread-major classes (KA), entering classes (NT), class records (CLS_LANES * 512)
                const int wl = tid >> 5;
                if ((tid & 31) == 0) {
                    constexpr int NWc = NT / 32;
                    if (wl == 4 % NWc) bulk_prefetch_l2_range(J.cperm + (size_t)(g + 1) * KA, (size_t)2 * KA);
                    if (wl == 5 % NWc) bulk_prefetch_l2_range(J.ccls + (size_t)(g + 1) * KA, (size_t)KA);
                    if (wl == 6 % NWc) bulk_prefetch_l2_range(J.cent + (size_t)(g + 1) * NT, (size_t)NT);
                    if (wl == 7 % NWc) bulk_prefetch_l2_range(J.crec + (size_t)(g + 1) * (CLS_LANES * 32), (size_t)CLS_LANES * 32 * 16);
                }
            }
        }
        QB_T(1);
        double cnew[NH];
        if (g == 0) {
            // rcpp_reinitialize_in_iterations
            double sv[NH];
#pragma unroll
            for (int h = 0; h < NH; h++) {
#pragma unroll
                for (int i = 0; i < EPT; i++) {
                    const int k = tid + i * NT;
                    am[h][i] = (k < K) ? one_over_K * eg[h * KA + k] : 0.0;
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
            QB_T(20);
            const double x = tm_x, t1 = tm_t1;
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
                        if (has) v = eg[h * KA + k] * v;
                    }
                    am[h][i] = v;
                }
                sv[h] = Col<NT, EPT>::sum(am[h]);
            }
            bsum.run(sv);
            QB_T(21);
#pragma unroll
            for (int h = 0; h < NH; h++) {
                double sc = 1 / (c_old[h] * sv[h]);
                cnew[h] = c_old[h] * sc;
                sc *= c_old[h];
#pragma unroll
                for (int i = 0; i < EPT; i++) am[h][i] *= sc;
            }
        }
        // am now holds alphaHat_t[:, g]; it stays in registers as the previous column of the next grid
        QB_T(2);
        // class path: worth its per-grid set-up (class totals, column factors) only with enough reads to visit.  The sorted
        // haplotype list, the classes of this thread's own elements and the records of this thread's classes are L2 hits
        // (prefetched one grid ahead), consumed after ab_m has been formed.
        constexpr int CPT = (CLS_LANES * 32 + NT - 1) / NT;  // classes per thread: class tid + j * NT
        uint32_t pq[EPT / 2];  // sorted positions [EPT * tid, EPT * tid + EPT): haplotype | first-of-its-class << 15
        uint32_t cq[EPT / 4];  // class of element i of this thread in byte i
        uint32_t cwp[CPT], cw0[CPT], cwn[CPT];  // allele words of this thread's classes (previous / own / next grid, masked)
        int ent1 = 0;
        uint64_t vis = 0;      // bit ir: read ir of the grid is visited (uninformative reads never are, gibbs-nipt.cpp:815)
        if (CLS_ON && use_cls) {
            const ReadDesc* dsc = reinterpret_cast<const ReadDesc*>(smem + L.off_small[s] + L.small_desc);
            const int lane = tid & 31;
            const bool va = (lane < fcn) && (reinterpret_cast<const uint4*>(dsc + lane)->y & 0xff) != 1;
            const bool vb = (32 + lane < fcn) && (reinterpret_cast<const uint4*>(dsc + 32 + lane)->y & 0xff) != 1;
            vis = (uint64_t)__ballot_sync(0xffffffffu, va) | ((uint64_t)__ballot_sync(0xffffffffu, vb) << 32);
            use_cls = __popcll(vis) >= P.cls_min_reads;
        }
        if (CLS_ON && use_cls) {
            const uint4* gp = reinterpret_cast<const uint4*>(J.cperm + (size_t)g * KA + (size_t)EPT * tid);
#pragma unroll
            for (int q = 0; q < EPT / 8; q++) {
                const uint4 v = __ldg(gp + q);
                pq[4 * q] = v.x, pq[4 * q + 1] = v.y, pq[4 * q + 2] = v.z, pq[4 * q + 3] = v.w;
            }
            const uint2* gc = reinterpret_cast<const uint2*>(J.ccls + (size_t)g * KA + (size_t)EPT * tid);
#pragma unroll
            for (int q = 0; q < EPT / 8; q++) {
                const uint2 v = __ldg(gc + q);
                cq[2 * q] = v.x, cq[2 * q + 1] = v.y;
            }
            ent1 = __ldg(J.cent + (size_t)g * NT + tid);
            const uint4* gr = J.crec + (size_t)g * (CLS_LANES * 32) + tid;
#pragma unroll
            for (int j = 0; j < CPT; j++) {
                const uint4 v = (tid + j * NT < CLS_LANES * 32) ? __ldg(gr + NT * j) : make_uint4(0u, 0u, 0u, 0u);
                cwp[j] = v.x, cw0[j] = v.y, cwn[j] = v.z;
            }
        }
        QB_T(14);
        if (g + 1 < T && (P.dbg & 2)) issue_pkg(g + 1, r1, nx_r1, ts1, nx_fcn, nx_fct);  // experiment: issue after the forward step
        bool changed = false;
        if (has) {
            const unsigned char* sm = smem + L.off_small[s];
            const ReadDesc* descs = reinterpret_cast<const ReadDesc*>(sm + L.small_desc);
            const TabEnt* tabs = reinterpret_cast<const TabEnt*>(sm + L.small_tab);
            const double* Us = reinterpret_cast<const double*>(sm + L.small_U);
            const int32_t* Hs = reinterpret_cast<const int32_t*>(sm + L.small_H);
            const P3 prior = {P.prior[0], P.prior[1], P.prior[2]};
            P3 pC = {1, 1, 1};
            bool inited = false;

            auto kind_of = [&](int r) -> int {
                // which of the three regimes (gibbs-nipt.cpp:816-834)
                if (iterative) {
                    if (iteration == 0) return (r < J.first_read) ? KIND_PASS : KIND_INIT;
                    if (iteration == 1 && r < J.first_read) return KIND_INIT;
                }
                return KIND_NORMAL;
            };
            auto init_ab = [&]() {
                // alphaHat_m = alpha, betaHat_m = beta, ab_m = alpha * beta ; pC = colsums (gibbs-nipt.cpp:836-858)
                double sv[NH];
                mbar_wait(&bar[5], n_useB & 1);
                n_useB++;
                beta_pending = false;
                double* bs = eGs + (size_t)((s ^ 1) * NH) * KA;
#pragma unroll
                for (int h = 0; h < NH; h++) {
#pragma unroll
                    for (int i = 0; i < EPT; i++) {
                        const int k = tid + i * NT;
                        ab[h][i] = (k < K) ? am[h][i] * bs[h * KA + k] : 0.0;
                    }
                    sv[h] = Col<NT, EPT>::sum(ab[h]);
                }
                if (CLS_ON && use_cls) {
                    // class path: ab_m takes the place of beta in the ring buffer (each thread overwrites what it has just read),
                    // to be gathered in class order; the buffer moves on to eMatGrid[:, g + 1] after the gather
#pragma unroll
                    for (int h = 0; h < NH; h++)
#pragma unroll
                        for (int i = 0; i < EPT; i++) bs[h * KA + tid + i * NT] = ab[h][i];
                    fence_proxy_async();
                }
                bsum.run(sv);  // (every thread is past its reads of the beta buffer)
                pC.a = sv[0];
                pC.b = sv[1];
                if (NH == 3) pC.c = sv[NH - 1];
                inited = true;
                if (!eG_next_issued && !(CLS_ON && use_cls)) {
                    issue_eG(g + 1);
                    eG_next_issued = true;
                }
            };
            // ESrc of chunk-local read ir; returns the source kind (0..3)
            auto make_src = [&](int ir, uint32_t tab0, ESrc& S) -> int {
                // descriptor as one 16-byte word (no local-memory copy): off | cat,mode,nb,g0rel | b0,sel0..2 | ...
                const uint4 dq = *reinterpret_cast<const uint4*>(descs + ir);
                const int mode = (dq.y >> 8) & 0xff, nb = (dq.y >> 16) & 0xff;
                const int g0rel = (int)(int8_t)(dq.y >> 24);
                S.b0 = dq.z & 0xff;
                S.mask = (1u << nb) - 1u;
                S.wlo = Wr + ((g + g0rel) & 3) * KA;
                S.whi = Wr + ((g + g0rel + 1) & 3) * KA;
                S.cross = S.b0 + nb > 32;
                S.tab = tabs + (dq.x - tab0);
                S.spat = spat;
                S.dcol = J.dense + (size_t)dq.x * Kp + kbase;
                if (mode == MODE_DENSE) return 3;
                if (mode == MODE_GATHER) return 2;
                return S.cross ? 1 : 0;
            };
            // divide the old label's alphaHat_m / ab_m / eMatGrid by the read's column, multiply the new label's
            // (gibbs-nipt.cpp:1092-1110).  Divisions: q = a * (1/e) corrected by one fma residual step = the correctly
            // rounded a / e (DESIGN.md "arithmetic"); dense columns divide.
#define QB_UPD_LOOP(SRC, HC, HN, DODIV) upd_loop<NT, EPT, NH, SRC, HC, HN, DODIV>(am, ab, eg, KA, S, K, tid);
#define QB_UPD_LABELS(SRC, normal, hC, hN)                                                                \
    if (normal) {                                                                                         \
        if (hC == 0 && hN == 1) {                                                                         \
            QB_UPD_LOOP(SRC, 0, 1, true)                                                                  \
        } else if (hC == 1 && hN == 0) {                                                                  \
            QB_UPD_LOOP(SRC, 1, 0, true)                                                                  \
        } else if (NH == 3) {                                                                             \
            if (hC == 0 && hN == 2) {                                                                     \
                QB_UPD_LOOP(SRC, 0, NH - 1, true)                                                         \
            } else if (hC == 1 && hN == 2) {                                                              \
                QB_UPD_LOOP(SRC, 1, NH - 1, true)                                                         \
            } else if (hC == 2 && hN == 0) {                                                              \
                QB_UPD_LOOP(SRC, NH - 1, 0, true)                                                         \
            } else if (hC == 2 && hN == 1) {                                                              \
                QB_UPD_LOOP(SRC, NH - 1, 1, true)                                                         \
            }                                                                                             \
        }                                                                                                 \
    } else {                                                                                              \
        if (hN == 0) {                                                                                    \
            QB_UPD_LOOP(SRC, 0, 0, false)                                                                 \
        } else if (hN == 1) {                                                                             \
            QB_UPD_LOOP(SRC, 0, 1, false)                                                                 \
        } else if (NH == 3) {                                                                             \
            QB_UPD_LOOP(SRC, 0, NH - 1, false)                                                            \
        }                                                                                                 \
    }
#define QB_UPD(src, normal, hC, hN)                                                                       \
    if (src == 0) {                                                                                       \
        QB_UPD_LABELS(0, normal, hC, hN)                                                                  \
    } else if (src == 1) {                                                                                \
        QB_UPD_LABELS(1, normal, hC, hN)                                                                  \
    } else if (src == 2) {                                                                                \
        QB_UPD_LABELS(2, normal, hC, hN)                                                                  \
    } else {                                                                                              \
        QB_UPD_LABELS(3, normal, hC, hN)                                                                  \
    }
            // K-long sums: normal -> sum ab_C / e, sum ab_A1 * e, (sum ab_A2 * e) ; init -> every label times e
#define QB_SUM_LOOP(SRC, XC, XA1, XA2, CMUL)                                                              \
    {                                                                                                     \
        double t0_ = 0, t1_ = 0, t2_ = 0; /* second accumulator set: halves the dependent fma chains */  \
        _Pragma("unroll") for (int i = 0; i < EPT; i++) {                                                 \
            const int k = tid + i * NT;                                                                   \
            if (SRC != 3 || k < K) {                                                                      \
                const EV ev = emission_at<SRC>(S, k);                                                     \
                if (i & 1) {                                                                       \
                    t0_ = fma(XC[i], CMUL ? ev.E : ev.invE, t0_);                                         \
                    t1_ = fma(XA1[i], ev.E, t1_);                                                         \
                    if (NH == 3) t2_ = fma(XA2[i], ev.E, t2_);                                            \
                } else {                                                                                  \
                    s0 = fma(XC[i], CMUL ? ev.E : ev.invE, s0);                                           \
                    s1 = fma(XA1[i], ev.E, s1);                                                           \
                    if (NH == 3) s2 = fma(XA2[i], ev.E, s2);                                              \
                }                                                                                         \
            }                                                                                             \
        }                                                                                                 \
        s0 += t0_;                                                                                        \
        s1 += t1_;                                                                                        \
        if (NH == 3) s2 += t2_;                                                                           \
    }
#define QB_SUM_LABELS(SRC, normal, hC)                                                                    \
    if (!(normal)) {                                                                                      \
        QB_SUM_LOOP(SRC, ab[0], ab[1], ab[NH - 1], true)                                                  \
    } else if (hC == 0) {                                                                                 \
        QB_SUM_LOOP(SRC, ab[0], ab[1], ab[NH - 1], false)                                                 \
    } else if (hC == 1) {                                                                                 \
        QB_SUM_LOOP(SRC, ab[1], ab[0], ab[NH - 1], false)                                                 \
    } else {                                                                                              \
        QB_SUM_LOOP(SRC, ab[NH - 1], ab[0], ab[1], false)                                                 \
    }
#define QB_SUM(src, normal, hC)                                                                           \
    if (src == 0) {                                                                                       \
        QB_SUM_LABELS(0, normal, hC)                                                                      \
    } else if (src == 1) {                                                                                \
        QB_SUM_LABELS(1, normal, hC)                                                                      \
    } else if (src == 2) {                                                                                \
        QB_SUM_LABELS(2, normal, hC)                                                                      \
    } else {                                                                                              \
        QB_SUM_LABELS(3, normal, hC)                                                                      \
    }

            // ---------------------------------------------------------------- one read at a time (any mode / regime)
            // ir = index inside the staged chunk, r = read index, tab0 = table-pool offset of the chunk
            auto single = [&](int ir, int r, uint32_t tab0) {
                const int kind = kind_of(r);
                if (!inited) init_ab();
                ESrc S;
                const int src = make_src(ir, tab0, S);
                if (src == 2 && kind != KIND_PASS) {
                    // gather mode: materialise the patterns once (each thread writes and later reads only its own elements)
                    const uint4 dq = *reinterpret_cast<const uint4*>(descs + ir);
                    const int nb = (dq.y >> 16) & 0xff;
                    const uint8_t* sel = reinterpret_cast<const uint8_t*>(descs + ir) + 9;
#pragma unroll
                    for (int i = 0; i < EPT; i++) {
                        const int k = tid + i * NT;
                        uint32_t pt = 0;
                        for (int j = 0; j < nb; j++) {
                            const int sj = sel[j];
                            pt |= ((Wr[((g + (sj >> 5) - 1) & 3) * KA + k] >> (sj & 31)) & 1u) << j;
                        }
                        spat[k] = (uint16_t)pt;
                    }
                }
                int hC = 0;
                double sv[NH];
#pragma unroll
                for (int h = 0; h < NH; h++) sv[h] = 0;
                if (kind != KIND_PASS) {
                    const bool normal = kind == KIND_NORMAL;
                    if (normal) hC = Hs[ir] - 1;
                    double s0 = 0, s1 = 0, s2 = 0;
                    QB_SUM(src, normal, hC)
                    sv[0] = s0;
                    sv[1] = s1;
                    if (NH == 3) sv[NH - 1] = s2;
                    bsum.run(sv);
                }
                const Decision D = decide_read<NH>(pC, sv, hC, kind, Us[ir], prior);
                pC = D.pCnew;
                if (D.change) {
                    changed = true;
                    if (tid == 0) J.H[r] = D.hN + 1;
                    const int hN = D.hN;
                    const bool normal = kind == KIND_NORMAL;
                    QB_UPD(src, normal, hC, hN)
                }
                if (record && tid == 0) {
                    J.xprob[4 * (size_t)r + 0] = D.x.a;
                    J.xprob[4 * (size_t)r + 1] = D.x.b;
                    J.xprob[4 * (size_t)r + 2] = D.x.c;
                    J.xprob[4 * (size_t)r + 3] = -1.0;  // already normalised
                }
            };

            bool cls_done = false;
            if constexpr (CLS_ON) {
            if (use_cls) {
                cls_done = true;
                // ---------------------------------------------------------------- class path (classes.cuh)
                // The grid's reads only distinguish the nC <= CLS_MAX haplotype classes of the grid, so the two K-long sums of a
                // read are sums over class totals of ab_m: thread t keeps the totals of class t (both labels) and the factor its
                // class has accumulated; per read one table look-up, two products and the block sum.  alphaHat_m / eMatGrid
                // receive the accumulated factor of their class once, at the end of the grid.
                const int lane = tid & 31, warp = tid >> 5;
                constexpr int NW = NT / 32;
                double* As = reinterpret_cast<double*>(smem + L.off_cls);  // [CLS_AS][2]: {label 0, label 1} of class c at index c + 1
                double* wsc = As + 2 * CLS_AS;                             // [NW][4] carries of the segmented scan
                double A0[CPT], A1[CPT], M0[CPT], M1[CPT];
                QB_T(11);
                init_ab();
                QB_T(3);
                // ---- class totals of ab_m: thread t owns the sorted positions [EPT t, EPT t + EPT); runs of one class inside a
                //      thread are added in order, runs that span threads meet in a segmented scan (fixed tree)
                {
                    const double* abS = eGs + (size_t)((s ^ 1) * NH) * KA;
                    double v0[EPT], v1[EPT];
#pragma unroll
                    for (int j = 0; j < EPT; j++) {
                        const int k = (int)((pq[j >> 1] >> ((j & 1) * 16)) & 0x7fffu);
                        v0[j] = abS[k];
                        v1[j] = abS[KA + k];
                    }
                    double rs0 = 0, rs1 = 0, P0 = 0, P1 = 0;
                    bool seen = false;
                    int c1 = ent1;
#pragma unroll
                    for (int j = 0; j < EPT; j++) {
                        const bool head = ((pq[j >> 1] >> ((j & 1) * 16 + 15)) & 1u) != 0;
                        if (head) {
                            if (!seen) {
                                P0 = rs0;  // the class that entered the thread ends here
                                P1 = rs1;
                                seen = true;
                            } else {
                                *reinterpret_cast<double2*>(As + 2 * c1) = make_double2(rs0, rs1);  // a class that lies inside the thread
                            }
                            c1++;
                            rs0 = 0;
                            rs1 = 0;
                        }
                        rs0 += v0[j];
                        rs1 += v1[j];
                    }
                    // (x, f): the sum the thread hands on — its last run if a class starts inside it, else its whole range
                    double x0 = rs0, x1 = rs1;
                    bool f = seen;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const double y0 = __shfl_up_sync(0xffffffffu, x0, d), y1 = __shfl_up_sync(0xffffffffu, x1, d);
                        const int gf = __shfl_up_sync(0xffffffffu, (int)f, d);
                        if (lane >= d && !f) {
                            x0 = y0 + x0;
                            x1 = y1 + x1;
                            f = gf != 0;
                        }
                    }
                    if (NW > 1) {
                        if (lane == 31) {
                            wsc[warp * 4] = x0;
                            wsc[warp * 4 + 1] = x1;
                            wsc[warp * 4 + 2] = f ? 1.0 : 0.0;
                        }
                        __syncthreads();
                    }
                    double w0 = 0, w1 = 0;
#pragma unroll
                    for (int w = 0; w < NW - 1; w++) {
                        if (w < warp) {
                            const double a = wsc[w * 4], b = wsc[w * 4 + 1];
                            const bool fw = wsc[w * 4 + 2] != 0.0;
                            w0 = fw ? a : w0 + a;
                            w1 = fw ? b : w1 + b;
                        }
                    }
                    const double p0 = __shfl_up_sync(0xffffffffu, x0, 1), p1 = __shfl_up_sync(0xffffffffu, x1, 1);
                    const int pf = __shfl_up_sync(0xffffffffu, (int)f, 1);
                    double cin0 = w0, cin1 = w1;
                    if (lane > 0) {
                        cin0 = pf ? p0 : w0 + p0;
                        cin1 = pf ? p1 : w1 + p1;
                    }
                    if (seen) *reinterpret_cast<double2*>(As + 2 * ent1) = make_double2(cin0 + P0, cin1 + P1);
                    if (tid == NT - 1) *reinterpret_cast<double2*>(As + 2 * c1) = make_double2(seen ? rs0 : cin0 + rs0, seen ? rs1 : cin1 + rs1);
                }
                __syncthreads();  // class totals complete; every thread is past its gather from the ring buffer
                if (!eG_next_issued) {
                    issue_eG(g + 1);
                    eG_next_issued = true;
                }
#pragma unroll
                for (int j = 0; j < CPT; j++) {
                    const int c = tid + NT * j;
                    const double2 a = (c < nC) ? *reinterpret_cast<const double2*>(As + 2 * (c + 1)) : make_double2(0.0, 0.0);
                    A0[j] = a.x;
                    A1[j] = a.y;
                    M0[j] = 1.0;
                    M1[j] = 1.0;
                }
                // (the exchange buffer carries the column factors at the end of the grid: every thread writes the slots it has just
                //  read, nobody else touches them in between)
                QB_T(12);
                // ---- the grid's reads; the emission values of the next visited read are fetched while the current one is decided
#define QB_CLS_LOAD(IR, E_, I_, H_, U_)                                                                  \
    {                                                                                                    \
        const uint4 dq_ = *reinterpret_cast<const uint4*>(descs + (IR));                                 \
        const int nb_ = (dq_.y >> 16) & 0xff;                                                            \
        const int g0_ = (int)(int8_t)(dq_.y >> 24);                                                      \
        const uint32_t b0_ = dq_.z & 0xff;                                                               \
        const uint32_t mask_ = (1u << nb_) - 1u;                                                         \
        const TabEnt* tab_ = tabs + (dq_.x - (uint32_t)ts0);                                             \
        _Pragma("unroll") for (int j = 0; j < CPT; j++) {                                                \
            const uint32_t lo_ = g0_ < 0 ? cwp[j] : (g0_ == 0 ? cw0[j] : cwn[j]);                        \
            const uint32_t hi_ = g0_ < 0 ? cw0[j] : cwn[j];                                              \
            const uint32_t pat_ = __funnelshift_r(lo_, hi_, b0_) & mask_;                                \
            const double2 te_ = *reinterpret_cast<const double2*>(tab_ + pat_);                          \
            E_[j] = te_.x;                                                                               \
            I_[j] = te_.y;                                                                               \
        }                                                                                                \
        H_ = Hs[(IR)] - 1;                                                                               \
        U_ = Us[(IR)];                                                                                   \
    }
                int ir = __ffsll((long long)vis) - 1;
                vis &= vis - 1;
                double Ev[CPT], Iv[CPT], uC;
                int hC;
                QB_CLS_LOAD(ir, Ev, Iv, hC, uC)
                while (true) {
                    const int r = r0 + ir;
                    QB_T(11);
                    QB_N(16);
                    double sv[NH];
                    {
                        double s0 = 0, s1 = 0;
#pragma unroll
                        for (int j = 0; j < CPT; j++) {
                            s0 = fma(hC == 0 ? A0[j] : A1[j], Iv[j], s0);
                            s1 = fma(hC == 0 ? A1[j] : A0[j], Ev[j], s1);
                        }
                        sv[0] = s0;
                        sv[1] = s1;
                    }
                    // the next visited read's static data (independent of this read's outcome)
                    const bool has_next = vis != 0;
                    const int ir2 = has_next ? __ffsll((long long)vis) - 1 : ir;
                    vis &= vis - 1;
                    double En[CPT], In[CPT], uN;
                    int hN2;
                    QB_CLS_LOAD(ir2, En, In, hN2, uN)
                    QB_T(4);
                    bsum.run(sv);
                    QB_T(5);
                    const FastDecision F = decide_diploid_fast(pC, sv[0], sv[1], hC, uC, prior);
                    bool change;
                    if (F.decided) {
                        change = F.hN != hC;
                        if (change) {
                            pC.a = (hC == 0) ? sv[0] : sv[1];
                            pC.b = (hC == 0) ? sv[1] : sv[0];
                        }
                        if (record && tid >= NT - 4) {
                            const int q = tid - (NT - 4);
                            const double v = q == 0 ? F.prod_pC : (q == 1 ? F.prod_pA1 : (q == 2 ? F.prod_pA2 : (double)hC));
                            J.xprob[4 * (size_t)r + q] = v;  // raw products of a read whose label was hC
                        }
                    } else {
                        const Decision D = decide_read<NH>(pC, sv, hC, KIND_NORMAL, uC, prior);
                        change = D.change;
                        pC = D.pCnew;
                        if (record && tid == 0) {
                            J.xprob[4 * (size_t)r + 0] = D.x.a;
                            J.xprob[4 * (size_t)r + 1] = D.x.b;
                            J.xprob[4 * (size_t)r + 2] = D.x.c;
                            J.xprob[4 * (size_t)r + 3] = -1.0;
                        }
                    }
                    QB_T(6);
                    if (change) {
                        QB_N(17);
                        changed = true;
                        if (tid == 0) J.H[r] = 2 - hC;  // the other label (1-based)
                        // gibbs-nipt.cpp:1092-1110 on the class totals and on the accumulated column factors
                        if (hC == 0) {
#pragma unroll
                            for (int j = 0; j < CPT; j++) {
                                A0[j] = div_by(A0[j], Ev[j], Iv[j]);
                                M0[j] = div_by(M0[j], Ev[j], Iv[j]);
                                A1[j] *= Ev[j];
                                M1[j] *= Ev[j];
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < CPT; j++) {
                                A1[j] = div_by(A1[j], Ev[j], Iv[j]);
                                M1[j] = div_by(M1[j], Ev[j], Iv[j]);
                                A0[j] *= Ev[j];
                                M0[j] *= Ev[j];
                            }
                        }
                        QB_T(7);
                    }
                    if (!has_next) break;
                    ir = ir2;
                    hC = hN2;
                    uC = uN;
#pragma unroll
                    for (int j = 0; j < CPT; j++) {
                        Ev[j] = En[j];
                        Iv[j] = In[j];
                    }
                }
#undef QB_CLS_LOAD
                QB_T(15);
                if (changed) {
                    // every element of alphaHat_m / eMatGrid takes the accumulated factor of its class
                    QB_T(11);
#pragma unroll
                    for (int j = 0; j < CPT; j++) {
                        const int c = tid + NT * j;
                        if (c < CLS_LANES * 32) *reinterpret_cast<double2*>(As + 2 * (c + 1)) = make_double2(M0[j], M1[j]);
                    }
                    __syncthreads();
#pragma unroll
                    for (int i = 0; i < EPT; i++) {
                        const int c = (int)((cq[i >> 2] >> ((i & 3) * 8)) & 0xffu);
                        const double2 m = *reinterpret_cast<const double2*>(As + 2 * (c + 1));
                        const int k = tid + i * NT;
                        am[0][i] *= m.x;
                        am[1][i] *= m.y;
                        eg[k] *= m.x;
                        eg[KA + k] *= m.y;
                    }
                    QB_T(13);
                }
            }
            }
            if (!cls_done) {
            // ---------------------------------------------------------------- the grid's reads, chunk by chunk
            const bool special_its = iterative && iteration <= 1;  // sweeps with pass-through / initialisation reads
            int c0 = 0;
            uint32_t tab0 = (uint32_t)ts0;
            int cn = fcn;                    // the first chunk arrived with the grid's package
            uint32_t tend = (uint32_t)fct;
            while (c0 < n_g) {
                if (c0 > 0) {
                    // over-full grid: the next chunk (host-computed boundaries, kept in the descriptor of its first read)
                    // is staged synchronously
                    __syncthreads();  // every thread is done with the previous chunk's staged data
                    const uint2 ch = *reinterpret_cast<const uint2*>(reinterpret_cast<const unsigned char*>(J.desc + r0 + c0) + 24);
                    cn = (int)(ch.x & 0xffffu);
                    tend = ch.y;
                    stage_small(s, r0 + c0, cn, (int)tab0, (int)(tend - tab0), false);
                }
                for (int ir = 0; ir < cn; ir++) {
                    const uint4 dq = *reinterpret_cast<const uint4*>(descs + ir);
                    if (NH == 2 && (dq.y & 0xff) == 1) continue;  // diploid: uninformative reads are never visited (gibbs-nipt.cpp:815)
                    const int r = r0 + c0 + ir;
                    const int mode = (dq.y >> 8) & 0xff;
                    if (NH != 2 || mode != MODE_RUN || (special_its && kind_of(r) != KIND_NORMAL)) {
                        single(ir, r, tab0);
                        continue;
                    }
                    // ---- hot path: diploid, table-mode read on consecutive SNPs, normal regime
                    QB_T(11);
                    if (!inited) init_ab();
                    QB_T(3);
                    QB_N(16);
                    ESrc S;
                    {
                        const int nb = (dq.y >> 16) & 0xff;
                        const int g0rel = (int)(int8_t)(dq.y >> 24);
                        S.b0 = dq.z & 0xff;
                        S.mask = (1u << nb) - 1u;
                        S.wlo = Wr + ((g + g0rel) & 3) * KA;
                        S.whi = Wr + ((g + g0rel + 1) & 3) * KA;
                        S.cross = S.b0 + nb > 32;
                        S.tab = tabs + (dq.x - tab0);
                    }
                    const int hC = Hs[ir] - 1;
                    double sv[NH];
                    {
                        double s0 = 0, s1 = 0, s2 = 0;
                        if (!S.cross) {
                            if (hC == 0) {
                                QB_SUM_LOOP(0, ab[0], ab[1], ab[NH - 1], false)
                            } else {
                                QB_SUM_LOOP(0, ab[1], ab[0], ab[NH - 1], false)
                            }
                        } else {
                            if (hC == 0) {
                                QB_SUM_LOOP(1, ab[0], ab[1], ab[NH - 1], false)
                            } else {
                                QB_SUM_LOOP(1, ab[1], ab[0], ab[NH - 1], false)
                            }
                        }
                        sv[0] = s0;
                        sv[1] = s1;
                        (void)s2;
                    }
                    QB_T(4);
                    bsum.run(sv);
                    QB_T(5);
                    const FastDecision F = decide_diploid_fast(pC, sv[0], sv[1], hC, Us[ir], prior);
                    Decision D;
                    if (F.decided) {
                        D.hN = F.hN;
                        D.change = F.hN != hC;
                        if (D.change) {
                            // hN is the other label: the column sums become the two K-long sums just formed
                            pC.a = (hC == 0) ? sv[0] : sv[1];
                            pC.b = (hC == 0) ? sv[1] : sv[0];
                        }
                        if (record && tid >= NT - 4) {
                            // (last warp, one lane per slot: warp 0 already carries the label store and the bulk-copy issue)
                            const int q = tid - (NT - 4);
                            const double v = q == 0 ? F.prod_pC : (q == 1 ? F.prod_pA1 : (q == 2 ? F.prod_pA2 : (double)hC));
                            J.xprob[4 * (size_t)r + q] = v;  // raw products of a read whose label was hC
                        }
                    } else {
                        D = decide_read<NH>(pC, sv, hC, KIND_NORMAL, Us[ir], prior);
                        pC = D.pCnew;
                        if (record && tid == 0) {
                            J.xprob[4 * (size_t)r + 0] = D.x.a;
                            J.xprob[4 * (size_t)r + 1] = D.x.b;
                            J.xprob[4 * (size_t)r + 2] = D.x.c;
                            J.xprob[4 * (size_t)r + 3] = -1.0;
                        }
                    }
                    QB_T(6);
                    if (D.change) {
                        QB_N(17);
                        changed = true;
                        if (tid == 0) J.H[r] = D.hN + 1;
                        if (!S.cross) {
                            if (hC == 0) {
                                QB_UPD_LOOP(0, 0, 1, true)
                            } else {
                                QB_UPD_LOOP(0, 1, 0, true)
                            }
                        } else {
                            if (hC == 0) {
                                QB_UPD_LOOP(1, 0, 1, true)
                            } else {
                                QB_UPD_LOOP(1, 1, 0, true)
                            }
                        }
                        QB_T(7);
                    }
                }
                tab0 = tend;
                c0 += cn;
            }
            }  // K-long path
#undef QB_SUM
#undef QB_SUM_LABELS
#undef QB_SUM_LOOP
#undef QB_UPD
#undef QB_UPD_LABELS
#undef QB_UPD_LOOP
            QB_T(11);
            QB_N(18);
            if (beta_pending) {
                // no read of this grid was visited: consume the beta copy so that its buffer can move on
                mbar_wait(&bar[5], n_useB & 1);
                n_useB++;
                beta_pending = false;
            }
            if (!eG_next_issued) {
                issue_eG(g + 1);
                eG_next_issued = true;
            }
            if (changed) {
                // gibbs-nipt.cpp:1262-1292: renormalise every haplotype's column and fold into c
                double sv[NH];
#pragma unroll
                for (int h = 0; h < NH; h++) sv[h] = Col<NT, EPT>::sum(am[h]);
                fence_proxy_async();  // this thread's updates of the eMatGrid column (generic proxy) before the bulk store below (async proxy)
                bsum.run(sv);         // (its barrier: every thread's updates are done and fenced)
#pragma unroll
                for (int h = 0; h < NH; h++) {
                    const double alphaConst = 1 / sv[h];
                    cnew[h] *= alphaConst;
#pragma unroll
                    for (int i = 0; i < EPT; i++) am[h][i] *= alphaConst;
                }
            }
            QB_T(19);
        }
#pragma unroll
        for (int h = 0; h < NH; h++) {
            // alphaHat_t in HBM is only read by the passes that may follow a sweep (gamma -> hapProbs after a sampling
            // sweep, the NIPT block episode, debug export); the next sweep rebuilds alpha from its registers
            if (store_alpha) Col<NT, EPT>::store(am[h], alphaG + ((size_t)h * T + g) * Kp, K);
            if (tid == 0) cG[h * T + g] = cnew[h];
            cfin[h] = cnew[h];
        }
        if (changed && tid == 0) {
            // the changed eMatGrid column already sits in the ring buffer: one bulk store per haplotype instead of EPT shared-memory
            // loads + EPT global stores per thread (the padding k in [K, Kp) travels along; nothing reads it)
#pragma unroll
            for (int h = 0; h < NH; h++) bulk_s2g(eGg + ((size_t)h * T + g) * Kp, eg + (size_t)h * KA, Kpl * 8);
            bulk_commit();
        }
        QB_T(8);
    }
    QB_T(11);

    // =============================================================== backward (Rcpp_run_backward_haploid_QUILT_faster)
    // the eMatGrid columns changed above left by bulk stores issued by thread 0: their group completes before the block barrier
    // behind which the bulk loads below start
    if (tid == 0) bulk_wait_all();  // the bulk stores of the changed eMatGrid columns are complete
    __threadfence();
    fence_proxy_async_all();
    __syncthreads();
    {
        double b[NH][EPT];
#pragma unroll
        for (int h = 0; h < NH; h++) {
#pragma unroll
            for (int i = 0; i < EPT; i++) b[h][i] = (tid + i * NT < K) ? cfin[h] : 0.0;
            Col<NT, EPT>::store(b[h], betaG + ((size_t)h * T + (T - 1)) * Kp, K);
        }
        // Stage ring for the backward walk: step g needs eMatGrid[:, g + 1], c[g] and the transition pair g.  The
        // allele-word ring is idle now and serves as a third column stage (diploid), so packages run NSB - 1 steps
        // ahead of their use; the step's scalars ride along by cp.async.
        constexpr int NSB = (NH == 2) ? 3 : 2;
        auto bstage = [&](int q) -> double* { return q < 2 ? eGs + (size_t)(q * NH) * KA : reinterpret_cast<double*>(Wr); };
        auto issue_b = [&](int g) {
            const int q = ((T - 2) - g) % NSB;
            if (tid == 0) {
                mbar_arrive_expect_tx(&bar[q], NH * Kpl * 8);
                double* dst = bstage(q);
#pragma unroll
                for (int h = 0; h < NH; h++) bulk_g2s(dst + (size_t)h * KA, eGg + ((size_t)h * T + g + 1) * Kp, Kpl * 8, &bar[q]);
            }
            unsigned char* sc = smem + L.off_sc + 192 + q * 64;
            const int ql = (NT > 32) ? tid - 32 : tid;
            if (ql >= 0 && ql < 8 + 2 * NH && (ql < 2 || ql >= 4)) {
                const int32_t* src;
                if (ql < 2)
                    src = rs + g + 1 + ql;
                else if (ql < 8)
                    src = reinterpret_cast<const int32_t*>(tmG + 2 * g) + (ql - 4);
                else
                    src = reinterpret_cast<const int32_t*>(cG + ((ql - 8) >> 1) * T + g) + ((ql - 8) & 1);
                cp_async4(sc + 4 * ql, src);
            }
        };
#pragma unroll
        for (int j = 0; j < NSB - 1; j++) {
            if (T - 2 - j >= 0) issue_b(T - 2 - j);
            cp_async_commit();
        }
        for (int g = T - 2; g >= 0; g--) {
            const int q = ((T - 2) - g) % NSB;
            // refill the stage step g + 1 used: every thread is past that step's block sum, i.e. done reading it
            if (g - (NSB - 1) >= 0) issue_b(g - (NSB - 1));
            cp_async_commit();
            cp_async_wait_group<NSB - 1>();
            if (q == 0) {
                mbar_wait(&bar[0], n_use0 & 1);
                n_use0++;
            } else if (q == 1) {
                mbar_wait(&bar[1], n_use1 & 1);
                n_use1++;
            } else {
                mbar_wait(&bar[2], n_use2 & 1);
                n_use2++;
            }
            __syncthreads();  // the other threads' cp.async scalars are visible
            const int32_t* sci = reinterpret_cast<const int32_t*>(smem + L.off_sc + 192 + q * 64);
            const double* scd = reinterpret_cast<const double*>(sci);
            const bool has1 = sci[1] > sci[0];
            const double* eg = bstage(q);
            const double t0 = scd[2], t1 = scd[3];
            double cg[NH], sv[NH];
#pragma unroll
            for (int h = 0; h < NH; h++) {
                cg[h] = scd[4 + h];
                if (has1) {
#pragma unroll
                    for (int i = 0; i < EPT; i++) {
                        const int k = tid + i * NT;
                        if (k < K) b[h][i] = eg[h * KA + k] * b[h][i];
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
                Col<NT, EPT>::store(b[h], betaG + ((size_t)h * T + g) * Kp, K);
            }
        }
    }

    // =============================================================== epilogue: H_class, counts, -sum log c, underflow
    __syncthreads();
    QB_T(9);
    for (int r = tid; r < R; r += NT) {
        const int h = J.H[r];
        atomicAdd(&cnt[h - 1], 1);
        if (record) {
            int hc;
            if (NH == 2 && J.desc[r].cat == 1) {
                hc = J.Hclass[r];
            } else {
                const double* xp = J.xprob + 4 * (size_t)r;
                double x0 = xp[0], x1 = xp[1], x2 = xp[2];
                if (xp[3] >= 0) {
                    // raw products of the fast decision path: normalise exactly as decide_read does
                    const double denom = x0 + x1 + x2;
                    const double nC = x0 / denom, nA1 = x1 / denom, nA2 = x2 / denom;
                    const bool c0 = xp[3] == 0.0;
                    x0 = c0 ? nC : nA1;
                    x1 = c0 ? nA1 : nC;
                    x2 = nA2;
                }
                hc = classify_H(P, x0, x1, x2);
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
        for (int g = tid + (int)crank * NT; g < T; g += NT * CL) {  // (cluster: the grids are split over the two CTAs, the block sum adds both)
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
#if QB_CLK
    QB_T(10);
    if (tid == QB_CLK_TID && blockIdx.x == 0 && iteration == QB_CLK_IT) {
        printf("QBCLK it=%d T=%d R=%d wait_pkg=%lld issue=%lld fwd=%lld init_ab=%lld sums=%lld reduce=%lld decide=%lld update=%lld gridend=%lld backward=%lld epilogue=%lld other=%lld "
               "class_sums=%lld class_apply=%lld visited=%lld changed=%lld grids_with_reads=%lld | cls_layout_issue=%lld reads_tail=%lld renorm=%lld fwd_sum1=%lld fwd_step=%lld\n",
               iteration, T, R, clk[0], clk[1], clk[2], clk[3], clk[4], clk[5], clk[6], clk[7], clk[8], clk[9], clk[10], clk[11], clk[12], clk[13], clk[16], clk[17], clk[18],
               clk[14], clk[15], clk[19], clk[20], clk[21]);
    }
#endif
}

}  // namespace qb
