// prep.cuh — per-call preparation kernels: panel bit-unpack and the per-read emission build.
//
//  k_unpack_common / k_assemble_all / k_scatter_rare / k_snp_type
//      the 32-SNP allele words of the K selected haplotypes (reference: rcpp_int_expand / inflate_fhb,
//      copied-from-stitch.cpp:50-108; compressed-panel lookup gibbs-small.cpp:204-231, :579-597;
//      special-haplotype search rcpp_simple_binary_matrix_search, gibbs-small.cpp:69-105)
//  k_build_tables / k_build_dense
//      eMatRead_t (Rcpp_make_eMatRead_t_for_gibbs_using_objects, gibbs-small.cpp:116-265, rare/common
//      sibling :275-465) followed by rcpp_evaluate_read_variability (gibbs-nipt.cpp:338-382)
//  k_expand_eMatRead
//      the dense K x R matrix from the tables (parity tests / QUILT_F_RETURN_EXTRA only)
#pragma once

#include "device_common.cuh"
#include "types.h"

namespace qb {

// gibbs-small.cpp:69-105, including its quirks: a single-row group returns 0 (not the stored word), the
// probe count is capped at 100 and the fallback returns row s1 (1-based s1 used as a 0-based row).
__device__ __forceinline__ int special_search(int val, const int32_t* __restrict__ mat, int nrow, int s1, int e1) {
    const int nori = e1 - s1 + 1;
    if (nori == 1) return 0;
    int n = nori;
    int i = n / 2;
    n = n / 4;
    int c = 0;
    while (c < 100) {
        c++;
        const int v = __ldg(mat + (s1 - 1 + i));
        if (v == val) return __ldg(mat + nrow + (s1 - 1 + i));
        if (v < val)
            i += n;
        else
            i -= n;
        n = n / 2;
        if (n < 1) n = 1;
        if (i < 0) i = 0;
        if (i > (nori - 1)) i = nori - 1;
    }
    return __ldg(mat + nrow + s1);
}

__device__ __forceinline__ uint32_t panel_word(const PanelDev& P, int k0, int g) {
    const int kk = __ldg(P.hapMatcherR + (size_t)g * P.K_full + k0);
    if (kk > 0) return (uint32_t)__ldg(P.distinctHapsB + (size_t)g * P.nMaxDH + (kk - 1));
    return (uint32_t)special_search(k0, P.special, P.n_special, __ldg(P.helper + g), __ldg(P.helper + P.Tc + g));
}

// words on the panel's (common-SNP) axis: out[g][k], grid = (ceil(Kp / 256), Tc, jobs)
__global__ void __launch_bounds__(256) k_unpack_common(PanelDev P, const JobDev* __restrict__ jobs, int K, int Kp, int to_Wc) {
    const JobDev& J = jobs[blockIdx.z];
    uint32_t* out = to_Wc ? J.Wc : J.W;
    const int g = blockIdx.y;
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (k >= Kp) return;
    uint32_t w = 0;
    if (k < K) w = panel_word(P, __ldg(J.which + k) - 1, g);
    out[(size_t)g * Kp + k] = w;
}

// all-SNP axis: bits of the common SNPs come from the common-axis words (rare_common.R:229-230), rare bits are
// or-ed in afterwards by k_scatter_rare.  Where the 32 SNPs of an all-SNP grid sit on the common axis is panel-only
// information: it is decoded ONCE per panel at upload (PanelDev::asm_src / asm_cg0: source bit inside one of the TWO
// common-axis words an all-SNP grid can touch — 32 consecutive SNPs hold at most 32 consecutive common SNPs), so a thread
// does two coalesced word loads and 32 shift / mask steps per grid, no dependent index loads and no barrier.
// grid = (ceil(Kp / 256), ceil(T_all / ASM_GPB), jobs): a CTA column of 256 haplotypes walks ASM_GPB consecutive grids.
constexpr int ASM_GPB = 8;
__global__ void __launch_bounds__(256) k_assemble_all(PanelDev P, const JobDev* __restrict__ jobs, int K, int Kp, int T_all) {
    const JobDev& J = jobs[blockIdx.z];
    const uint32_t* __restrict__ Wc = J.Wc;
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (k >= Kp) return;
    const int G0 = blockIdx.y * ASM_GPB, G1 = min(G0 + ASM_GPB, T_all);
    for (int G = G0; G < G1; G++) {
        uint32_t w = 0;
        if (k < K) {
            const int cg0 = __ldg(P.asm_cg0 + G);
            const uint32_t c0 = Wc[(size_t)cg0 * Kp + k];
            const uint32_t c1 = (cg0 + 1 < P.Tc) ? Wc[(size_t)(cg0 + 1) * Kp + k] : 0u;
            const int8_t* __restrict__ src = P.asm_src + 32 * (size_t)G;
#pragma unroll
            for (int b4 = 0; b4 < 32; b4 += 4) {
                const char4 s4 = *reinterpret_cast<const char4*>(src + b4);
                const int sv[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (sv[q] >= 0) w |= ((((sv[q] >> 5) ? c1 : c0) >> (sv[q] & 31)) & 1u) << (b4 + q);
            }
        }
        J.W[(size_t)G * Kp + k] = w;
    }
}

// rare_per_hap_info (rare_common.R:202-322): grid = (ceil(K / 256), jobs)
__global__ void __launch_bounds__(256) k_scatter_rare(PanelDev P, const JobDev* __restrict__ jobs, int K, int Kp) {
    const JobDev& J = jobs[blockIdx.y];
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (k >= K) return;
    const int h = __ldg(J.which + k) - 1;
    for (int64_t j = P.rare_off[h]; j < P.rare_off[h + 1]; j++) {
        const int s = __ldg(P.rare_snps + j) - 1;
        J.W[(size_t)(s >> 5) * Kp + k] |= 1u << (s & 31);
    }
}

// snp_type on the all-SNP axis: 0 common, 1 rare without a carrier among the selected haplotypes
// ("k_with_alt.length() == 1", gibbs-small.cpp:404), 2 rare with carrier(s).  grid = (T_all, jobs), 128 threads
__global__ void __launch_bounds__(128) k_snp_type(PanelDev P, const JobDev* __restrict__ jobs, int K, int Kp) {
    const JobDev& J = jobs[blockIdx.y];
    const int G = blockIdx.x;
    __shared__ uint32_t mask;
    if (threadIdx.x == 0) mask = 0;
    __syncthreads();
    uint32_t m = 0;
    for (int k = threadIdx.x; k < K; k += 128) m |= J.W[(size_t)G * Kp + k];
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) m |= __shfl_xor_sync(0xffffffffu, m, d);
    if ((threadIdx.x & 31) == 0) atomicOr(&mask, m);
    __syncthreads();
    if (threadIdx.x < 32) {
        const int s = 32 * G + threadIdx.x;
        if (s < P.nSNPs_all) {
            uint8_t t = 0;
            if (!P.snp_is_common[s]) t = ((mask >> threadIdx.x) & 1u) ? 2 : 1;
            J.snp_type[s] = t;
        }
    }
}

// allele pattern of haplotype k over the read's SNPs, words read from global memory (build-time path)
__device__ __forceinline__ uint32_t read_pattern_global(const ReadDesc& d, const uint32_t* __restrict__ W, int Kp, int g, int k) {
    if (d.mode == MODE_RUN) {
        const uint32_t lo = W[(size_t)(g + d.g0rel) * Kp + k];
        const uint32_t hi = (d.b0 + d.nb > 32) ? W[(size_t)(g + d.g0rel + 1) * Kp + k] : 0u;
        return __funnelshift_r(lo, hi, d.b0) & ((1u << d.nb) - 1u);
    }
    uint32_t pat = 0;
    for (int j = 0; j < d.nb; j++) {
        const int wr = d.sel[j] >> 5, b = d.sel[j] & 31;
        pat |= ((W[(size_t)(g + wr - 1) * Kp + k] >> b) & 1u) << j;
    }
    return pat;
}

// one factor of the per-read emission product for a haplotype whose allele at the SNP is `bit`
// common SNP: gibbs-small.cpp:204-231 ; rare SNP: gibbs-small.cpp:399-428
__device__ __forceinline__ double emission_factor_apply(double E, int type, bool rescale, uint32_t bit, double pR, double pA, double eps) {
    if (type == 0) {
        const double e = bit ? (1 - eps) : eps;
        return E * (e * pA + (1 - e) * pR);
    }
    if (type == 1 && rescale) return E;  // no carrier: skipped entirely when rescaling (gibbs-small.cpp:404-411)
    const double ome = 1 - eps;
    const double xe1 = eps * pA + ome * pR;
    E = E * xe1;
    if (bit) {
        const double xe2 = ome * pA + eps * pR;
        E = E * (xe2 / xe1);
    }
    return E;
}

constexpr int TAB_WARPS = 8;

// allele pattern from the three staged word columns (grids g-1, g, g+1) in shared memory
__device__ __forceinline__ uint32_t read_pattern_staged(const ReadDesc& d, const uint32_t* Ws, int Kp, int k) {
    if (d.mode == MODE_RUN) {
        const uint32_t lo = Ws[(d.g0rel + 1) * Kp + k];
        const uint32_t hi = (d.b0 + d.nb > 32) ? Ws[(d.g0rel + 2) * Kp + k] : 0u;
        return __funnelshift_r(lo, hi, d.b0) & ((1u << d.nb) - 1u);
    }
    uint32_t pat = 0;
    for (int j = 0; j < d.nb; j++) {
        const int wr = d.sel[j] >> 5, b = d.sel[j] & 31;
        pat |= ((Ws[wr * Kp + k] >> b) & 1u) << j;
    }
    return pat;
}

// One CTA per (grid, job): the allele words a read of this grid can touch (grids g-1 .. g+1) are staged in shared
// memory once, then each warp takes table-mode reads of the grid in turn.  grid = (T, jobs), 256 threads,
// dynamic shared memory = 3 * Kp * 4 bytes.
__global__ void __launch_bounds__(TAB_WARPS * 32) k_build_tables(BatchParams P, const JobDev* __restrict__ jobs) {
    extern __shared__ __align__(16) uint32_t Ws[];  // [3][Kp]
    __shared__ int hist[TAB_WARPS][1 << NBMAX];
    const JobDev& J = jobs[blockIdx.y];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.x, Kp = P.Kp, T = P.T;
    const int r0 = J.rs[g], r1 = J.rs[g + 1];
    if (r0 >= r1) return;
    const int nC = J.cinfo ? J.cinfo[g] : 0;  // > 0: every read of the grid is a table-mode run and the class records exist
    if (nC == 0)
    for (int i = threadIdx.x; i < 3 * Kp; i += TAB_WARPS * 32) {
        const int w = i / Kp, k = i - w * Kp;
        const int gg = g - 1 + w;
        Ws[i] = (gg >= 0 && gg < T) ? J.W[(size_t)gg * Kp + k] : 0u;
    }
    __syncthreads();
    const bool rescale = (P.flags & QUILT_F_RESCALE_EMATREAD) != 0;
    int* hs = hist[warp];
    for (int r = r0 + warp; r < r1; r += TAB_WARPS) {
    ReadDesc d = J.desc[r];
    if (d.mode == MODE_DENSE) continue;
    const int nb = d.nb, n = 1 << nb;
    const int uo = J.roff[r];
    TabEnt* tab = J.tabs + d.off;
    __syncwarp();
    // 1. raw products per allele pattern, factors applied in read order
    for (int pat = lane; pat < n; pat += 32) {
        double E = 1.0;
        for (int j = 0; j < nb; j++) {
            const int s = J.u[uo + j];
            const double pR = J.pRA[2 * (size_t)(uo + j)], pA = J.pRA[2 * (size_t)(uo + j) + 1];
            const int type = P.rare_common ? J.snp_type[s] : 0;
            E = emission_factor_apply(E, type, rescale, (pat >> j) & 1u, pR, pA, P.ref_error);
        }
        tab[pat].E = E;
        hs[pat] = 0;
    }
    __syncwarp();
    // 2. how many of the K haplotypes show each pattern
    if (nC > 0) {
        // the grid has haplotype classes (classes.cuh): members of a class show the same pattern — add the class sizes
        const uint4* rec = J.crec + (size_t)g * (CLS_LANES * 32);
        for (int c = lane; c < nC; c += 32) {
            const uint4 v = rec[c];
            const uint32_t lo = d.g0rel < 0 ? v.x : (d.g0rel == 0 ? v.y : v.z), hi = d.g0rel < 0 ? v.y : v.z;
            const uint32_t pat = __funnelshift_r(lo, hi, d.b0) & ((1u << d.nb) - 1u);
            atomicAdd(&hs[pat], (int)(v.w >> 16));
        }
        __syncwarp();
    } else
    // (lanes showing the same pattern elect one leader that adds their count: no shared-memory atomics)
    // (four independent match.any per trip: its latency is what bounds this loop; peer masks from ten ballots measured 55 % slower)
    for (int k0 = 0; k0 < P.K; k0 += 128) {
        uint32_t pat[4], peers[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int k = k0 + 32 * q + lane;
            pat[q] = (k < P.K) ? read_pattern_staged(d, Ws, Kp, k) : 0xffffffffu;
        }
#pragma unroll
        for (int q = 0; q < 4; q++) peers[q] = __match_any_sync(0xffffffffu, pat[q]);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (pat[q] != 0xffffffffu && lane == __ffs(peers[q]) - 1) hs[pat[q]] += __popc(peers[q]);
            __syncwarp();
        }
    }
    // 3. rescale by the maximum over the haplotypes present, then floor (gibbs-small.cpp:235-262)
    bool degenerate = false;
    double d1 = 1.0;
    if (rescale) {
        double x = 0;
        for (int pat = lane; pat < n; pat += 32) {
            const double E = tab[pat].E;
            if (hs[pat] > 0 && E > x) x = E;
        }
#pragma unroll
        for (int dd = 16; dd >= 1; dd >>= 1) {
            const double o = __shfl_xor_sync(0xffffffffu, x, dd);
            if (o > x) x = o;
        }
        d1 = 1 / x;
        const double inf = __longlong_as_double(0x7ff0000000000000LL);
        degenerate = (x == inf) | (x == 0) | (d1 == inf) | (d1 == -inf);
    }
    // 4. category (gibbs-nipt.cpp:338-382) on the rescaled values of the patterns present
    int nn = 0;
    double vmin = __longlong_as_double(0x7ff0000000000000LL), vmax = -vmin;
    for (int pat = lane; pat < n; pat += 32) {
        double E = tab[pat].E;
        if (rescale) {
            if (degenerate) {
                E = 1;
            } else {
                E *= d1;
                if (E < P.d2) E = P.d2;
            }
        }
        tab[pat].E = E;
        tab[pat].invE = 1 / E;
        if (hs[pat] > 0 && E < ONE_THRESH) {
            nn += hs[pat];
            vmin = fmin(vmin, E);
            vmax = fmax(vmax, E);
        }
    }
#pragma unroll
    for (int dd = 16; dd >= 1; dd >>= 1) {
        nn += __shfl_xor_sync(0xffffffffu, nn, dd);
        vmin = fmin(vmin, __shfl_xor_sync(0xffffffffu, vmin, dd));
        vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, dd));
    }
    if (lane == 0) {
        const int thresh2 = (int)(P.K * 0.20);
        int cat;
        if (nn == 0)
            cat = 1;
        else if (vmin == vmax)
            cat = 2;
        else if (nn < thresh2)
            cat = 3;
        else
            cat = 0;
        if ((P.flags & QUILT_F_FORCE_RESET_READ_CATEGORY_0) && cat != 1) cat = 0;
        if (P.flags & QUILT_F_DISABLE_READ_CATEGORY_USAGE) cat = 0;
        J.desc[r].cat = (uint8_t)cat;
    }
    }  // reads of the grid
}

// one CTA per dense-mode read (many SNPs, or SNPs outside grids wif0-1..wif0+1): the K-long column itself.
// grid = (max n_dense, jobs), 256 threads; dense_reads[job] lists the read indices
__global__ void __launch_bounds__(256) k_build_dense(BatchParams P, const JobDev* __restrict__ jobs) {
    const JobDev& J = jobs[blockIdx.y];
    if ((int)blockIdx.x >= J.n_dense) return;
    const int r = J.dense_reads[blockIdx.x];
    const ReadDesc d = J.desc[r];
    TabEnt* col = J.dense + (size_t)d.off * P.Kp;
    const int uo = J.roff[r];
    int cnt = J.roff[r + 1] - uo;
    if (cnt - 1 >= P.Jmax) cnt = P.Jmax + 1;
    const bool rescale = (P.flags & QUILT_F_RESCALE_EMATREAD) != 0;
    __shared__ double sred[8];
    __shared__ int sint[8], sfirst[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double x = 0;
    for (int k = threadIdx.x; k < P.Kp; k += 256) {
        double E = 1.0;
        if (k < P.K) {
            for (int j = 0; j < cnt; j++) {
                const int s = J.u[uo + j];
                const double pR = J.pRA[2 * (size_t)(uo + j)], pA = J.pRA[2 * (size_t)(uo + j) + 1];
                const int type = P.rare_common ? J.snp_type[s] : 0;
                const uint32_t bit = (J.W[(size_t)(s >> 5) * P.Kp + k] >> (s & 31)) & 1u;
                E = emission_factor_apply(E, type, rescale, bit, pR, pA, P.ref_error);
            }
            if (E > x) x = E;
        }
        col[k].E = E;
        col[k].invE = 1.0;
    }
    bool degenerate = false;
    double d1 = 1;
    if (rescale) {
#pragma unroll
        for (int dd = 16; dd >= 1; dd >>= 1) {
            const double o = __shfl_xor_sync(0xffffffffu, x, dd);
            if (o > x) x = o;
        }
        if (lane == 0) sred[warp] = x;
        __syncthreads();
        x = sred[0];
        for (int w = 1; w < 8; w++)
            if (sred[w] > x) x = sred[w];
        __syncthreads();
        d1 = 1 / x;
        const double inf = __longlong_as_double(0x7ff0000000000000LL);
        degenerate = (x == inf) | (x == 0) | (d1 == inf) | (d1 == -inf);
    }
    // rescale + find the first non-1 haplotype (smallest k) and the count
    int nn = 0, kfirst = 0x7fffffff;
    for (int k = threadIdx.x; k < P.K; k += 256) {
        double E = col[k].E;
        if (rescale) {
            if (degenerate) {
                E = 1;
            } else {
                E *= d1;
                if (E < P.d2) E = P.d2;
            }
            col[k].E = E;
        }
        col[k].invE = 1 / E;
        if (E < ONE_THRESH) {
            nn++;
            if (k < kfirst) kfirst = k;
        }
    }
#pragma unroll
    for (int dd = 16; dd >= 1; dd >>= 1) {
        nn += __shfl_xor_sync(0xffffffffu, nn, dd);
        kfirst = min(kfirst, __shfl_xor_sync(0xffffffffu, kfirst, dd));
    }
    if (lane == 0) {
        sint[warp] = nn;
        sfirst[warp] = kfirst;
    }
    __syncthreads();
    nn = 0;
    kfirst = 0x7fffffff;
    for (int w = 0; w < 8; w++) {
        nn += sint[w];
        kfirst = min(kfirst, sfirst[w]);
    }
    __syncthreads();
    int more = 0;
    if (nn > 0) {
        const double val = col[kfirst].E;
        for (int k = threadIdx.x; k < P.K; k += 256) {
            const double E = col[k].E;
            if (E < ONE_THRESH && E != val) more = 1;
        }
    }
    more = __syncthreads_or(more);
    if (threadIdx.x == 0) {
        const int thresh2 = (int)(P.K * 0.20);
        int cat;
        if (nn == 0)
            cat = 1;
        else if (!more)
            cat = 2;
        else if (nn < thresh2)
            cat = 3;
        else
            cat = 0;
        if ((P.flags & QUILT_F_FORCE_RESET_READ_CATEGORY_0) && cat != 1) cat = 0;
        if (P.flags & QUILT_F_DISABLE_READ_CATEGORY_USAGE) cat = 0;
        J.desc[r].cat = (uint8_t)cat;
    }
}

// dense eMatRead_t [K x R] (column-major, stride K) for the parity tests.  grid = (R, jobs = 1), 256 threads
__global__ void __launch_bounds__(256) k_expand_eMatRead(BatchParams P, const JobDev* __restrict__ jobs, double* __restrict__ out) {
    const JobDev& J = jobs[0];
    const int r = blockIdx.x;
    const ReadDesc d = J.desc[r];
    const int g = J.wif0[r];
    for (int k = threadIdx.x; k < P.K; k += 256) {
        double E;
        if (d.mode == MODE_DENSE)
            E = J.dense[(size_t)d.off * P.Kp + k].E;
        else
            E = J.tabs[d.off + read_pattern_global(d, J.W, P.Kp, g, k)].E;
        out[(size_t)r * P.K + k] = E;
    }
}

}  // namespace qb
