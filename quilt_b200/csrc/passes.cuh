// passes.cuh — the remaining K x T passes of one Gibbs call.
//
//  k_init_iterative   rcpp_gibbs_nipt_initialize with gibbs_initialize_iteratively (gibbs-nipt.cpp:1725-1740)
//  k_make_eG          rcpp_make_eMatGrid_t, bound = false (copied-from-stitch.cpp:234-310)
//  k_fb_generic       Rcpp_run_forward_haploid + Rcpp_run_backward_haploid with the uniform 1/K prior
//                     (copied-from-stitch.cpp:340-409; wrapper gibbs-nipt.cpp:453-487)
//  k_shard            Rcpp_shard_block_gibbs_resampler, ff == 0, shard_check_every_pair
//                     (gibbs-nipt-block.cpp:1975-2355; generic forward step gibbs-nipt.cpp:630-661)
//  k_happrobs         unpack_gammas -> rcpp_calculate_gibbs_small_genProbs_and_hapProbs_using_binary_objects
//                     (gibbs-small.cpp:472-635) and the rare/common sibling (:711-867), plus the
//                     equal-weight running average of rcpp_fly_weighter (gibbs-nipt.cpp:2043-2094)
#pragma once

#include "device_common.cuh"
#include "prep.cuh"
#include "sweep.cuh"
#include "types.h"

namespace qb {

// grid = (T, jobs), 256 threads
__global__ void __launch_bounds__(256) k_init_iterative(BatchParams P, const JobDev* __restrict__ jobs) {
    const JobDev& J = jobs[blockIdx.y];
    const int g = blockIdx.x, K = P.K, Kp = P.Kp, T = P.T, NH = P.NH;
    __shared__ double sred[8];
    double a0 = 1.0, c0 = 1.0;
    if (g == 0) {
        // alpha[:, 0] = prior * eMatGrid[:, 0] (= prior), c[0] = 1 / sum, alpha *= c[0]
        double s = 0;
        for (int k = threadIdx.x; k < K; k += 256) s += P.one_over_K * 1.0;
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = s;
        __syncthreads();
        s = sred[0];
        for (int w = 1; w < 8; w++) s += sred[w];
        c0 = 1 / s;
        a0 = P.one_over_K * c0;
    }
    for (int h = 0; h < NH; h++) {
        const size_t o = ((size_t)h * T + g) * Kp;
        for (int k = threadIdx.x; k < Kp; k += 256) {
            const bool in = k < K;
            J.eG[o + k] = 1.0;
            J.alpha[o + k] = in ? a0 : 0.0;
            J.beta[o + k] = in ? 1.0 : 0.0;
        }
        if (threadIdx.x == 0) J.c[h * T + g] = c0;
    }
}

// grid = (T, jobs), 256 threads: eMatGrid[:, g] of every haplotype = product of its reads' columns, in read order
// (rcpp_make_eMatGrid_t, bound = false).  The grid's descriptors, labels and emission tables are staged in shared memory in
// the HOST-computed chunks the sweep kernel uses too (ginfo row of the grid, chunk fields of the descriptors: no serial scan
// on the device); a thread works on four haplotypes at a time — their three allele words in registers, four independent
// product chains per haplotype label — so a table-mode factor costs a shift, a mask and one shared-memory load.
__global__ void __launch_bounds__(256) k_make_eG(BatchParams P, const JobDev* __restrict__ jobs) {
    __shared__ JobDev Js;
    __shared__ __align__(16) ReadDesc sdesc[SW_MAXR];
    __shared__ double stab[SW_MAXTAB];
    __shared__ int sH[SW_MAXR];
    const int tid = threadIdx.x;
    if (tid == 0) Js = jobs[blockIdx.y];
    __syncthreads();
    const JobDev& J = Js;
    const int g = blockIdx.x, K = P.K, Kp = P.Kp, T = P.T, NH = P.NH;
    const int r0 = J.ginfo[4 * g], r1 = J.rs[g + 1];
    const int n_g = r1 - r0;
    int c0 = 0, cn = J.ginfo[4 * g + 2];
    uint32_t tab0 = (uint32_t)J.ginfo[4 * g + 1], tend = (uint32_t)J.ginfo[4 * g + 3];
    bool first = true;
    constexpr int G4 = 4;
    do {
        if (c0 > 0) {
            // later chunk of an over-full grid: its size / table end ride in the descriptor of its first read
            const ReadDesc& d0 = J.desc[r0 + c0];
            cn = d0.chunk_n;
            tend = d0.chunk_tend;
        }
        __syncthreads();  // the previous chunk's staged data is no longer read
        for (int i = tid; i < cn * 2; i += 256) reinterpret_cast<uint4*>(sdesc)[i] = reinterpret_cast<const uint4*>(J.desc + r0 + c0)[i];
        for (int i = tid; i < cn; i += 256) sH[i] = J.H[r0 + c0 + i];
        for (int i = tid; i < (int)(tend - tab0); i += 256) stab[i] = J.tabs[tab0 + i].E;
        __syncthreads();
        for (int kb = tid; kb < Kp; kb += 256 * G4) {
            double e[G4][3];
            uint32_t wm[G4], w0[G4], wp[G4];
#pragma unroll
            for (int q = 0; q < G4; q++) {
                const int k = kb + q * 256;
                const bool in = k < Kp;
                for (int h = 0; h < 3; h++) e[q][h] = (first || !in || h >= NH) ? 1.0 : J.eG[((size_t)h * T + g) * Kp + k];
                const bool live = k < K && cn > 0;
                wm[q] = (live && g > 0) ? J.W[(size_t)(g - 1) * Kp + k] : 0u;
                w0[q] = live ? J.W[(size_t)g * Kp + k] : 0u;
                wp[q] = (live && g + 1 < T) ? J.W[(size_t)(g + 1) * Kp + k] : 0u;
            }
            // (padding haplotypes K <= k < Kp carry zero words; whatever factors they pick up are discarded at the store: padding stays 1)
            for (int ir = 0; ir < cn; ir++) {
                const uint4 dq = *reinterpret_cast<const uint4*>(sdesc + ir);
                const int mode = (dq.y >> 8) & 0xff, nb = (dq.y >> 16) & 0xff;
                const int h = sH[ir] - 1;
                double E[G4];
                // every branch below is uniform over the CTA (it depends on the read only), so the per-haplotype work is a shift,
                // a mask, one shared-memory load and one multiply
                if (mode == MODE_RUN) {
                    const int g0rel = (int)(int8_t)(dq.y >> 24);
                    const uint32_t b0 = dq.z & 0xff, mask = (1u << nb) - 1u;
                    const double* tb = stab + (dq.x - tab0);
                    if (b0 + nb <= 32) {
                        if (g0rel == 0) {
#pragma unroll
                            for (int q = 0; q < G4; q++) E[q] = tb[(w0[q] >> b0) & mask];
                        } else if (g0rel < 0) {
#pragma unroll
                            for (int q = 0; q < G4; q++) E[q] = tb[(wm[q] >> b0) & mask];
                        } else {
#pragma unroll
                            for (int q = 0; q < G4; q++) E[q] = tb[(wp[q] >> b0) & mask];
                        }
                    } else if (g0rel < 0) {
#pragma unroll
                        for (int q = 0; q < G4; q++) E[q] = tb[__funnelshift_r(wm[q], w0[q], b0) & mask];
                    } else {
#pragma unroll
                        for (int q = 0; q < G4; q++) E[q] = tb[__funnelshift_r(g0rel == 0 ? w0[q] : wp[q], wp[q], b0) & mask];
                    }
                } else if (mode == MODE_DENSE) {
#pragma unroll
                    for (int q = 0; q < G4; q++) {
                        const int k = kb + q * 256;
                        E[q] = (k < K) ? J.dense[(size_t)dq.x * Kp + k].E : 1.0;
                    }
                } else {
                    const uint8_t* sel = reinterpret_cast<const uint8_t*>(sdesc + ir) + 9;
                    const double* tb = stab + (dq.x - tab0);
#pragma unroll
                    for (int q = 0; q < G4; q++) {
                        uint32_t pat = 0;
                        for (int b = 0; b < nb; b++) {
                            const int wr = sel[b] >> 5, bit = sel[b] & 31;
                            const uint32_t w = wr == 0 ? wm[q] : (wr == 1 ? w0[q] : wp[q]);
                            pat |= ((w >> bit) & 1u) << b;
                        }
                        E[q] = tb[pat];
                    }
                }
                if (h == 0) {
#pragma unroll
                    for (int q = 0; q < G4; q++) e[q][0] *= E[q];
                } else if (h == 1) {
#pragma unroll
                    for (int q = 0; q < G4; q++) e[q][1] *= E[q];
                } else if (h == 2) {
#pragma unroll
                    for (int q = 0; q < G4; q++) e[q][2] *= E[q];
                }
            }
#pragma unroll
            for (int q = 0; q < G4; q++) {
                const int k = kb + q * 256;
                if (k < Kp)
                    for (int h = 0; h < NH; h++) J.eG[((size_t)h * T + g) * Kp + k] = (k < K) ? e[q][h] : 1.0;
            }
        }
        first = false;
        c0 += cn;
        tab0 = tend;
        if (cn == 0) break;  // (empty grid: one pass wrote the ones)
    } while (c0 < n_g);
}

// generic forward + backward of one haplotype.  grid = (jobs, NH).  If ext_* are given they replace the job's
// arrays (component entry point quilt_gpu_forward_backward).
template <int NT, int EPT>
__global__ void __launch_bounds__(NT) k_fb_generic(BatchParams P, const JobDev* __restrict__ jobs, int do_backward) {
    __shared__ double red[2 * SW_VMAX * (NT / 32)];
    const JobDev& J = jobs[blockIdx.x];
    if (*J.underflow) return;
    const int h = blockIdx.y, tid = threadIdx.x;
    const int K = P.K, Kp = P.Kp, T = P.T;
    BlockSumV<NT> bsum(red);
    const double prior = P.one_over_K;
    const double* eG = J.eG + (size_t)h * T * Kp;
    double* alpha = J.alpha + (size_t)h * T * Kp;
    double* beta = J.beta + (size_t)h * T * Kp;
    double* c = J.c + h * T;
    double a[EPT], e[EPT], en[EPT];
    Col<NT, EPT>::load(e, eG, K, 0.0);
    double clast = 1;
    for (int g = 0; g < T; g++) {
        if (g + 1 < T) Col<NT, EPT>::load(en, eG + (size_t)(g + 1) * Kp, K, 0.0);
        if (g == 0) {
#pragma unroll
            for (int i = 0; i < EPT; i++) a[i] = prior * e[i];
        } else {
            const double t0 = J.tm[2 * (g - 1)], t1 = J.tm[2 * (g - 1) + 1];
#pragma unroll
            for (int i = 0; i < EPT; i++) a[i] = (tid + i * NT < K) ? e[i] * (t0 * a[i] + t1 * prior) : 0.0;
        }
        double sv[1] = {Col<NT, EPT>::sum(a)};
        bsum.run(sv);
        const double cg = 1 / sv[0];
#pragma unroll
        for (int i = 0; i < EPT; i++) a[i] *= cg;
        Col<NT, EPT>::store(a, alpha + (size_t)g * Kp, K);
        if (tid == 0) c[g] = cg;
        clast = cg;
#pragma unroll
        for (int i = 0; i < EPT; i++) e[i] = en[i];
    }
    if (!do_backward) return;
    __syncthreads();
    double b[EPT];
#pragma unroll
    for (int i = 0; i < EPT; i++) b[i] = (tid + i * NT < K) ? clast : 0.0;
    Col<NT, EPT>::store(b, beta + (size_t)(T - 1) * Kp, K);
    if (T >= 2) Col<NT, EPT>::load(e, eG + (size_t)(T - 1) * Kp, K, 0.0);
    for (int g = T - 2; g >= 0; g--) {
        if (g >= 1) Col<NT, EPT>::load(en, eG + (size_t)g * Kp, K, 0.0);
        const double cg = ld_cg(c + g);
        const double t0 = J.tm[2 * g], t1 = J.tm[2 * g + 1];
        double sv[1] = {0};
#pragma unroll
        for (int i = 0; i < EPT; i++) {
            b[i] = e[i] * b[i];
            sv[0] += prior * b[i];
        }
        bsum.run(sv);
        const double x = t1 * sv[0];
#pragma unroll
        for (int i = 0; i < EPT; i++) b[i] = (tid + i * NT < K) ? cg * (x + t0 * b[i]) : 0.0;
        Col<NT, EPT>::store(b, beta + (size_t)g * Kp, K);
#pragma unroll
        for (int i = 0; i < EPT; i++) e[i] = en[i];
    }
}

// L2 prefetch of NC columns of Kp doubles starting at p (column stride `stride` doubles), spread over the CTA
template <int NT>
__device__ __forceinline__ void prefetch_cols(const double* p, size_t stride, int ncols, int Kp) {
    const int per = Kp >> 4;  // 128-byte lines per column
    for (int l = threadIdx.x; l < ncols * per; l += NT) {
        const int c = l / per, q = l - c * per;
        prefetch_l2(p + (size_t)c * stride + (q << 4));
    }
}

// shard pass (diploid).  grid = jobs (x CL).  One CTA (or two-CTA cluster, CL = 2: each CTA owns half of the K states,
// see k_sweep) re-runs the forward recursion of both haplotypes with the generic step and, after every grid, decides
// between "stay" and "swap labels from here on"; then the generic backward of both haplotypes in one walk.
//
// Columns arrive by bulk copy (TMA engine, mbarrier completion) in a ring of three column PAIRS in shared memory with
// rotating tenants: eMatGrid[:, g] sits in pair 2g mod 3, beta[:, g] in pair (2g + 1) mod 3.  When the forward step has
// consumed eMatGrid[:, g] its pair receives beta[:, g + 1]; when the scores have consumed beta[:, g] its pair receives
// eMatGrid[:, g + 2] — every column is in flight for at least a full step, nothing waits in registers.
// dynamic shared memory = 3 * 2 * KA * 8 bytes.
template <int NT, int EPT, int CL = 1>
__global__ void __launch_bounds__(NT) k_shard(BatchParams P, const JobDev* __restrict__ jobs, int episode) {
    extern __shared__ __align__(128) unsigned char shsm[];
    __shared__ double red[2 * CL * SW_VMAX * (NT / 32)];
    __shared__ JobDev Js;
    __shared__ __align__(8) uint64_t bar[3];
    constexpr int KA = NT * EPT;
    const int tid = threadIdx.x;
    if (tid == 0) Js = jobs[blockIdx.x / CL];
    __syncthreads();
    if (*Js.underflow) return;
    const uint32_t crank = (CL > 1) ? cluster_ctarank() : 0u;
    const int kbase = (CL > 1) ? (int)crank * KA : 0;
    const int K = (CL > 1) ? max(0, min(P.K - kbase, KA)) : P.K;       // states of this CTA
    const int Kp = P.Kp, T = P.T, R = Js.R;
    const int Kpl = (CL > 1) ? max(0, min(P.Kp - kbase, KA)) : P.Kp;   // padded states of this CTA (bulk-copy length)
    double* __restrict__ alphaG = Js.alpha + kbase;
    double* __restrict__ betaG = Js.beta + kbase;
    double* __restrict__ eGg = Js.eG + kbase;
    double* cG = Js.c;
    double* rateG = Js.rate;
    const double* __restrict__ tmG = Js.tm;
    BlockSumV<NT, CL> bsum(red, crank);
    const double prior = P.one_over_K;
    const double* __restrict__ runif = Js.runif_shard + (size_t)episode * (T - 1);
    const size_t hs = (size_t)T * Kp;  // haplotype stride
    double* ring = reinterpret_cast<double*>(shsm);  // [3 pairs][2 haplotypes][KA]
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_init(&bar[2], 1);
        fence_barrier_init();
    }
    __syncthreads();
    uint32_t use0 = 0, use1 = 0, use2 = 0;
    // both haplotypes' column g of `src` into ring pair q (caller: every thread is done with that pair)
    auto issue = [&](int q, const double* src, int g) {
        if (tid == 0) {
            fence_proxy_async();
            mbar_arrive_expect_tx(&bar[q], 2 * Kpl * 8);
            bulk_g2s(ring + (size_t)(q * 2) * KA, src + (size_t)g * Kp, Kpl * 8, &bar[q]);
            bulk_g2s(ring + (size_t)(q * 2 + 1) * KA, src + hs + (size_t)g * Kp, Kpl * 8, &bar[q]);
        }
    };
    auto wait = [&](int q) {
        if (q == 0) {
            mbar_wait(&bar[0], use0 & 1);
            use0++;
        } else if (q == 1) {
            mbar_wait(&bar[1], use1 & 1);
            use1++;
        } else {
            mbar_wait(&bar[2], use2 & 1);
            use2++;
        }
    };
    issue(0, eGg, 0);                 // eMatGrid[:, 0] -> pair 0
    if (T > 1) {
        issue(1, betaG, 0);           // beta[:, 0]     -> pair 1
        issue(2, eGg, 1);             // eMatGrid[:, 1] -> pair 2
    }
    double mloc[2];
    {
        double sl[2] = {0, 0};
        for (int g = tid + (int)crank * NT; g < T; g += NT * CL) {  // (cluster: grids split over the two CTAs)
            sl[0] += log(ld_cg(cG + g));
            sl[1] += log(ld_cg(cG + T + g));
        }
        bsum.run(sl);
        mloc[0] = -sl[0];
        mloc[1] = -sl[1];
    }
    double mlc[2] = {0, 0};
    bool in_flip = false;
    double ap[2][EPT];
    double clast[2] = {1, 1};
    double nx_c[2] = {ld_cg(cG), ld_cg(cG + T)};
    double nx_x = 0, nx_t1 = 0, nx_u = (T > 1) ? runif[0] : 0.0;
    int qe = 0, qb = 1;  // ring pairs of eMatGrid[:, g] and beta[:, g]
    for (int g = 0; g < T; g++) {
        const double orig_c[2] = {nx_c[0], nx_c[1]};
        const double x = nx_x, t1 = nx_t1, u = nx_u;
        if (g + 1 < T) {
            nx_c[0] = ld_cg(cG + g + 1);
            nx_c[1] = ld_cg(cG + T + g + 1);
            nx_x = tmG[2 * g];
            nx_t1 = tmG[2 * g + 1];
            if (g + 1 < T - 1) nx_u = runif[g + 1];
        }
        wait(qe);
        const double* e0 = ring + (size_t)(qe * 2 + (in_flip && g > 0 ? 1 : 0)) * KA;  // flip mode: the columns trade places
        const double* e1 = ring + (size_t)(qe * 2 + (in_flip && g > 0 ? 0 : 1)) * KA;
        double cn[2];
        if (g == 0) {
            double sv[2];
#pragma unroll
            for (int i = 0; i < EPT; i++) {
                const int k = tid + i * NT;
                ap[0][i] = (k < K) ? prior * e0[k] : 0.0;
                ap[1][i] = (k < K) ? prior * e1[k] : 0.0;
            }
            sv[0] = Col<NT, EPT>::sum(ap[0]);
            sv[1] = Col<NT, EPT>::sum(ap[1]);
            bsum.run(sv);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                cn[h] = 1 / sv[h];
#pragma unroll
                for (int i = 0; i < EPT; i++) ap[h][i] *= cn[h];
            }
        } else {
            if (in_flip) {
                // eMatGrid_t1.col(g) <-> eMatGrid_t2.col(g), in place
                double t0v[EPT], t1v[EPT];
#pragma unroll
                for (int i = 0; i < EPT; i++) {
                    const int k = tid + i * NT;
                    t0v[i] = (k < K) ? e0[k] : 0.0;
                    t1v[i] = (k < K) ? e1[k] : 0.0;
                }
                Col<NT, EPT>::store(t0v, eGg + (size_t)g * Kp, K);
                Col<NT, EPT>::store(t1v, eGg + hs + (size_t)g * Kp, K);
            }
            double sp[2];
#pragma unroll
            for (int h = 0; h < 2; h++) sp[h] = Col<NT, EPT>::sum(ap[h]);
            bsum.run(sp);
            double sv[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const double alphaConst = t1 * sp[h];
                const double jump = alphaConst * prior;
                const double c2 = orig_c[h];
                const double* eh = h == 0 ? e0 : e1;
#pragma unroll
                for (int i = 0; i < EPT; i++) {
                    const int k = tid + i * NT;
                    ap[h][i] = (k < K) ? (c2 * eh[k]) * (x * ap[h][i] + jump) : 0.0;
                }
                sv[h] = Col<NT, EPT>::sum(ap[h]);
            }
            bsum.run(sv);  // (every thread is past its reads of eMatGrid[:, g])
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const double sc = 1 / sv[h];
                cn[h] = orig_c[h] * sc;
#pragma unroll
                for (int i = 0; i < EPT; i++) ap[h][i] *= sc;
            }
        }
        // eMatGrid[:, g] is consumed (the block sums above are behind every thread's reads): its pair takes beta[:, g + 1]
        if (g + 1 < T - 1) issue(qe, betaG, g + 1);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            Col<NT, EPT>::store(ap[h], alphaG + h * hs + (size_t)g * Kp, K);
            if (tid == 0) cG[h * T + g] = cn[h];
            mlc[h] -= log(cn[h]);
            clast[h] = cn[h];
        }
        if (tid == 0) rateG[g] = in_flip ? 1.0 : 0.0;  // reads of this grid are relabelled 3 - H when set
        if (g < T - 1) {
            wait(qb);
            const double* y0 = ring + (size_t)(qb * 2) * KA;
            const double* y1 = y0 + KA;
            double dv[4] = {0, 0, 0, 0};
#pragma unroll
            for (int i = 0; i < EPT; i++) {
                const int k = tid + i * NT;
                const double b0 = (k < K) ? y0[k] : 0.0, b1 = (k < K) ? y1[k] : 0.0;
                dv[0] += ap[0][i] * b0;
                dv[1] += ap[1][i] * b1;
                dv[2] += ap[1][i] * b0;
                dv[3] += ap[0][i] * b1;
            }
            bsum.run(dv);  // (every thread is past its reads of beta[:, g])
            if (g + 2 < T) issue(qb, eGg, g + 2);
            const double pA1 = mlc[0] + mloc[0] + log(dv[0]);
            const double pA2 = mlc[1] + mloc[1] + log(dv[1]);
            const double pB1 = mlc[1] + mloc[0] + log(dv[2]);
            const double pB2 = mlc[0] + mloc[1] + log(dv[3]);
            const double diff = pB1 + pB2 - pA1 - pA2;
            double probs1 = 1;
            const double probs2 = exp(diff);
            const double psum = probs1 + probs2;
            probs1 /= psum;
            in_flip = u > probs1;
        }
#pragma unroll
        for (int h = 0; h < 2; h++) mloc[h] += log(orig_c[h]);
        // tenants of the next grid: eMatGrid[:, g + 1] sits in the pair that was neither e nor beta of this grid
        const int qn = 3 - qe - qb;
        qb = qe;
        qe = qn;
    }
    __syncthreads();
    // relabel the reads of every grid walked in flip mode (once per job: the first CTA of a cluster)
    if (crank == 0) {
        for (int r = tid; r < R; r += NT) {
            if (rateG[Js.wif0[r]] != 0.0) Js.H[r] = 3 - Js.H[r];
        }
    }
    // generic backward on the (possibly swapped) eMatGrid columns, both haplotypes in one walk.  The swapped columns
    // were written through the generic proxy by this CTA: order them before the bulk (async proxy) reads below.
    __threadfence();
    fence_proxy_async_all();
    __syncthreads();
    {
        double b[2][EPT];
#pragma unroll
        for (int h = 0; h < 2; h++) {
#pragma unroll
            for (int i = 0; i < EPT; i++) b[h][i] = (tid + i * NT < K) ? clast[h] : 0.0;
            Col<NT, EPT>::store(b[h], betaG + h * hs + (size_t)(T - 1) * Kp, K);
        }
        // step g needs eMatGrid[:, g + 1]; step index j = T - 2 - g uses pair j mod 3, packages run two steps ahead
        if (T >= 2) issue(0, eGg, T - 1);
        if (T >= 3) issue(1, eGg, T - 2);
        double nb_c[2] = {0, 0}, nb_t0 = 0, nb_t1 = 0;
        if (T >= 2) {
            nb_c[0] = ld_cg(cG + T - 2);
            nb_c[1] = ld_cg(cG + T + T - 2);
            nb_t0 = tmG[2 * (T - 2)];
            nb_t1 = tmG[2 * (T - 2) + 1];
        }
        int q = 0;
        for (int g = T - 2; g >= 0; g--) {
            const double cg[2] = {nb_c[0], nb_c[1]};
            const double t0 = nb_t0, t1 = nb_t1;
            if (g >= 1) {
                nb_c[0] = ld_cg(cG + g - 1);
                nb_c[1] = ld_cg(cG + T + g - 1);
                nb_t0 = tmG[2 * (g - 1)];
                nb_t1 = tmG[2 * (g - 1) + 1];
            }
            // the pair step g + 1 used is free (every thread passed that step's block sum): fill it for step g - 2
            if (g - 2 >= 0) issue((q + 2) % 3, eGg, g - 1);
            wait(q);
            const double* ee0 = ring + (size_t)(q * 2) * KA;
            const double* ee1 = ee0 + KA;
            double sv[2] = {0, 0};
#pragma unroll
            for (int i = 0; i < EPT; i++) {
                const int k = tid + i * NT;
                if (k < K) {
                    b[0][i] = ee0[k] * b[0][i];
                    b[1][i] = ee1[k] * b[1][i];
                }
                sv[0] += prior * b[0][i];
                sv[1] += prior * b[1][i];
            }
            bsum.run(sv);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const double x = t1 * sv[h];
#pragma unroll
                for (int i = 0; i < EPT; i++) b[h][i] = (tid + i * NT < K) ? cg[h] * (x + t0 * b[h][i]) : 0.0;
                Col<NT, EPT>::store(b[h], betaG + h * hs + (size_t)g * Kp, K);
            }
            q = (q + 1) % 3;
        }
    }
}

// gamma -> hapProbs / genProbs.  grid = (T, jobs), 128 * NH threads: each haplotype owns four warps; a thread keeps the
// 32 per-SNP alt sums of its k's in registers (k = j + 128 i), so gamma is formed once per (k, h) straight from the
// alpha / beta columns and never staged.  ref sums are total - alt.  The warp-level reduction is transposed (31
// shuffle-adds leave lane b with SNP b's total).
// first = this is the first sampling sweep (assign), otherwise accumulate; scale = 1 / n_sample applied on the last.
template <int NH>
__global__ void __launch_bounds__(128 * NH, NH == 2 ? 2 : 1) k_happrobs(BatchParams P, const JobDev* __restrict__ jobs, int first, int last, double scale) {
    __shared__ JobDev Js;
    __shared__ double part[3][4][33];
    const int tid = threadIdx.x;
    if (tid == 0) Js = jobs[blockIdx.y];
    __syncthreads();
    const JobDev& J = Js;
    if (*J.underflow) return;
    const int g = blockIdx.x, K = P.K, Kp = P.Kp, T = P.T, nSNPs = P.nSNPs;
    const int h = tid >> 7, j = tid & 127, lane = tid & 31, wq = (tid >> 5) & 3;
    double acc[32];
#pragma unroll
    for (int b = 0; b < 32; b++) acc[b] = 0.0;
    double tot = 0.0;
    {
        const double* __restrict__ a = J.alpha + ((size_t)h * T + g) * Kp;
        const double* __restrict__ bt = J.beta + ((size_t)h * T + g) * Kp;
        const uint32_t* __restrict__ W = J.W + (size_t)g * Kp;
        constexpr int U = 4;
        // register double-buffer: the next chunk's loads are in flight while this chunk's 32 x U adds run
        double av[2][U], bv[2][U];
        uint32_t wv[2][U];
        auto fetch = [&](int buf, int k0) {
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int k = k0 + 128 * u;
                const bool in = k < K;
                av[buf][u] = in ? ld_stream(a + k) : 0.0;
                bv[buf][u] = in ? ld_stream(bt + k) : 0.0;
                wv[buf][u] = in ? __ldg(W + k) : 0u;
            }
        };
        fetch(0, j);
        const double x = 1 / ld_cg(J.c + h * T + g);
        int cur = 0;
#pragma unroll 1
        for (int k0 = j; k0 < K; k0 += 128 * U * 2) {
            // two chunks per trip so that the buffer index is a compile-time constant in each half
            if (k0 + 128 * U < K) fetch(1, k0 + 128 * U);
#pragma unroll
            for (int u = 0; u < U; u++) {
                const double gk = (av[0][u] * bv[0][u]) * x;
                tot += gk;
#pragma unroll
                for (int b = 0; b < 32; b++)
                    if ((wv[0][u] >> b) & 1u) acc[b] += gk;
            }
            if (k0 + 128 * U < K) {
                if (k0 + 128 * U * 2 < K) fetch(0, k0 + 128 * U * 2);
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const double gk = (av[1][u] * bv[1][u]) * x;
                    tot += gk;
#pragma unroll
                    for (int b = 0; b < 32; b++)
                        if ((wv[1][u] >> b) & 1u) acc[b] += gk;
                }
            }
        }
        (void)cur;
    }
    warp_transpose_reduce<32>(acc, lane);  // lane b now holds SNP b's alt sum of this warp
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, d);
    part[h][wq][lane] = acc[0];
    if (lane == 0) part[h][wq][32] = tot;
    __syncthreads();
    if (tid < 32) {
        const int bit = tid;
        const int s = 32 * g + bit;
        if (s < nSNPs) {
            double A[3] = {0, 0, 0}, Rf[3] = {0, 0, 0};
#pragma unroll
            for (int hh = 0; hh < NH; hh++) {
                const double al = (part[hh][0][bit] + part[hh][1][bit]) + (part[hh][2][bit] + part[hh][3][bit]);
                const double tt = (part[hh][0][32] + part[hh][1][32]) + (part[hh][2][32] + part[hh][3][32]);
                A[hh] = al;
                Rf[hh] = tt - al;
            }
            const double eps = P.ref_error, ome = 1 - P.ref_error;
            double hp[3], gM[3], gF[3] = {0, 0, 0};
            if (!P.rare_common) {
                for (int hh = 0; hh < 3; hh++) hp[hh] = A[hh] * ome + Rf[hh] * eps;
                gM[0] = (1 - hp[0]) * (1 - hp[1]);
                gM[1] = (hp[0] * (1 - hp[1]) + (1 - hp[0]) * hp[1]);
                gM[2] = hp[0] * hp[1];
                gF[0] = (1 - hp[0]) * (1 - hp[2]);
                gF[1] = (hp[0] * (1 - hp[2]) + (1 - hp[0]) * hp[2]);
                gF[2] = hp[0] * hp[2];
            } else {
                // the reference accumulates into its (never re-zeroed) local matrix: hapLocal carries that state
                const int type = J.snp_type[s];
                for (int hh = 0; hh < 3; hh++) {
                    double v = J.hapLocal[hh * (size_t)nSNPs + s];
                    if (hh < NH) {
                        if (type == 0)
                            v += A[hh] * ome + Rf[hh] * eps;
                        else if (type == 1)
                            v = eps;
                        else
                            v += (A[hh] + Rf[hh]) * eps + A[hh] * (1 - 2 * eps);
                    }
                    hp[hh] = v;
                    J.hapLocal[hh * (size_t)nSNPs + s] = v;
                }
                gM[0] = (1 - hp[0]) * (1 - hp[1]);
                gM[1] = hp[0] * (1 - hp[1]) + hp[1] * (1 - hp[0]);
                gM[2] = hp[0] * hp[1];
                if (NH == 3) {
                    gF[0] = (1 - hp[0]) * (1 - hp[2]);
                    gF[1] = hp[0] * (1 - hp[2]) + hp[2] * (1 - hp[0]);
                    gF[2] = hp[0] * hp[2];
                }
            }
            for (int hh = 0; hh < 3; hh++) {
                // output layout [3 x nSNPs] column-major (row = haplotype / genotype)
                const size_t o = (size_t)s * 3 + hh;
                double vh = hp[hh], vm = gM[hh], vf = gF[hh];
                if (!first) {
                    vh += J.hapProbs[o];
                    vm += J.genM[o];
                    vf += J.genF[o];
                }
                if (last && scale != 1.0) {
                    vh *= scale;
                    vm *= scale;
                    vf *= scale;
                }
                J.hapProbs[o] = vh;
                J.genM[o] = vm;
                J.genF[o] = vf;
            }
        }
    }
}

}  // namespace qb
