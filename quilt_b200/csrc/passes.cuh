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
__global__ void __launch_bounds__(256) k_make_eG(BatchParams P, const JobDev* __restrict__ jobs) {
    const JobDev& J = jobs[blockIdx.y];
    const int g = blockIdx.x, K = P.K, Kp = P.Kp, T = P.T, NH = P.NH;
    const int r0 = J.rs[g], r1 = J.rs[g + 1];
    for (int k = threadIdx.x; k < Kp; k += 256) {
        double e[3] = {1.0, 1.0, 1.0};
        if (k < K) {
            for (int r = r0; r < r1; r++) {
                const ReadDesc d = J.desc[r];
                double E;
                if (d.mode == MODE_DENSE)
                    E = J.dense[(size_t)d.off * Kp + k];
                else
                    E = J.tabs[d.off + read_pattern_global(d, J.W, Kp, g, k)].E;
                const int h = J.H[r] - 1;
                if (h == 0)
                    e[0] *= E;
                else if (h == 1)
                    e[1] *= E;
                else if (h == 2)
                    e[2] *= E;
            }
        }
        for (int h = 0; h < NH; h++) J.eG[((size_t)h * T + g) * Kp + k] = e[h];
    }
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

// shard pass (diploid).  grid = jobs.  One CTA re-runs the forward recursion of both haplotypes with the
// generic step and, after every grid, decides between "stay" and "swap labels from here on".
template <int NT, int EPT>
__global__ void __launch_bounds__(NT) k_shard(BatchParams P, const JobDev* __restrict__ jobs, int episode) {
    __shared__ double red[2 * SW_VMAX * (NT / 32)];
    __shared__ JobDev Js;
    const int tid = threadIdx.x;
    if (tid == 0) Js = jobs[blockIdx.x];
    __syncthreads();
    const JobDev& J = Js;
    if (*J.underflow) return;
    const int K = P.K, Kp = P.Kp, T = P.T, R = J.R;
    BlockSumV<NT> bsum(red);
    const double prior = P.one_over_K;
    const double* __restrict__ runif = J.runif_shard + (size_t)episode * (T - 1);
    double mloc[2];
    {
        double sl[2] = {0, 0};
        for (int g = tid; g < T; g += NT) {
            sl[0] += log(ld_cg(J.c + g));
            sl[1] += log(ld_cg(J.c + T + g));
        }
        bsum.run(sl);
        mloc[0] = -sl[0];
        mloc[1] = -sl[1];
    }
    double mlc[2] = {0, 0};
    bool in_flip = false;
    double ap[2][EPT], e[2][EPT], en[2][EPT];
#pragma unroll
    for (int h = 0; h < 2; h++) Col<NT, EPT>::load(e[h], J.eG + ((size_t)h * T) * Kp, K, 0.0);
    double clast[2] = {1, 1};
    for (int g = 0; g < T; g++) {
        double orig_c[2];
#pragma unroll
        for (int h = 0; h < 2; h++) orig_c[h] = ld_cg(J.c + h * T + g);
        if (g + 1 < T) {
#pragma unroll
            for (int h = 0; h < 2; h++) Col<NT, EPT>::load(en[h], J.eG + ((size_t)h * T + g + 1) * Kp, K, 0.0);
        }
        double y[2][EPT];
        if (g < T - 1) {
#pragma unroll
            for (int h = 0; h < 2; h++) Col<NT, EPT>::load(y[h], J.beta + ((size_t)h * T + g) * Kp, K, 0.0);
        }
        double cn[2];
        if (g == 0) {
            double sv[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
#pragma unroll
                for (int i = 0; i < EPT; i++) ap[h][i] = prior * e[h][i];
                sv[h] = Col<NT, EPT>::sum(ap[h]);
            }
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
#pragma unroll
                for (int i = 0; i < EPT; i++) {
                    const double t = e[0][i];
                    e[0][i] = e[1][i];
                    e[1][i] = t;
                }
#pragma unroll
                for (int h = 0; h < 2; h++) Col<NT, EPT>::store(e[h], J.eG + ((size_t)h * T + g) * Kp, K);
            }
            double sp[2];
#pragma unroll
            for (int h = 0; h < 2; h++) sp[h] = Col<NT, EPT>::sum(ap[h]);
            bsum.run(sp);
            const double x = J.tm[2 * (g - 1)], t1 = J.tm[2 * (g - 1) + 1];
            double sv[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const double alphaConst = t1 * sp[h];
                const double jump = alphaConst * prior;
                const double c2 = orig_c[h];
#pragma unroll
                for (int i = 0; i < EPT; i++) ap[h][i] = (tid + i * NT < K) ? (c2 * e[h][i]) * (x * ap[h][i] + jump) : 0.0;
                sv[h] = Col<NT, EPT>::sum(ap[h]);
            }
            bsum.run(sv);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const double sc = 1 / sv[h];
                cn[h] = orig_c[h] * sc;
#pragma unroll
                for (int i = 0; i < EPT; i++) ap[h][i] *= sc;
            }
        }
#pragma unroll
        for (int h = 0; h < 2; h++) {
            Col<NT, EPT>::store(ap[h], J.alpha + ((size_t)h * T + g) * Kp, K);
            if (tid == 0) J.c[h * T + g] = cn[h];
            mlc[h] -= log(cn[h]);
            clast[h] = cn[h];
        }
        if (tid == 0) J.rate[g] = in_flip ? 1.0 : 0.0;  // reads of this grid are relabelled 3 - H when set
        if (g < T - 1) {
            double dv[4] = {0, 0, 0, 0};
#pragma unroll
            for (int i = 0; i < EPT; i++) {
                dv[0] += ap[0][i] * y[0][i];
                dv[1] += ap[1][i] * y[1][i];
                dv[2] += ap[1][i] * y[0][i];
                dv[3] += ap[0][i] * y[1][i];
            }
            bsum.run(dv);
            const double pA1 = mlc[0] + mloc[0] + log(dv[0]);
            const double pA2 = mlc[1] + mloc[1] + log(dv[1]);
            const double pB1 = mlc[1] + mloc[0] + log(dv[2]);
            const double pB2 = mlc[0] + mloc[1] + log(dv[3]);
            const double diff = pB1 + pB2 - pA1 - pA2;
            double probs1 = 1;
            const double probs2 = exp(diff);
            const double psum = probs1 + probs2;
            probs1 /= psum;
            in_flip = runif[g] > probs1;
        }
#pragma unroll
        for (int h = 0; h < 2; h++) {
            mloc[h] += log(orig_c[h]);
#pragma unroll
            for (int i = 0; i < EPT; i++) e[h][i] = en[h][i];
        }
    }
    __syncthreads();
    // relabel the reads of every grid walked in flip mode
    for (int r = tid; r < R; r += NT) {
        if (J.rate[J.wif0[r]] != 0.0) J.H[r] = 3 - J.H[r];
    }
    // generic backward on the (possibly swapped) eMatGrid columns
    for (int h = 0; h < 2; h++) {
        const double* eG = J.eG + (size_t)h * T * Kp;
        double* beta = J.beta + (size_t)h * T * Kp;
        double b[EPT], ee[EPT], een[EPT];
#pragma unroll
        for (int i = 0; i < EPT; i++) b[i] = (tid + i * NT < K) ? clast[h] : 0.0;
        Col<NT, EPT>::store(b, beta + (size_t)(T - 1) * Kp, K);
        if (T >= 2) Col<NT, EPT>::load(ee, eG + (size_t)(T - 1) * Kp, K, 0.0);
        for (int g = T - 2; g >= 0; g--) {
            if (g >= 1) Col<NT, EPT>::load(een, eG + (size_t)g * Kp, K, 0.0);
            const double cg = ld_cg(J.c + h * T + g);
            const double t0 = J.tm[2 * g], t1 = J.tm[2 * g + 1];
            double sv[1] = {0};
#pragma unroll
            for (int i = 0; i < EPT; i++) {
                b[i] = ee[i] * b[i];
                sv[0] += prior * b[i];
            }
            bsum.run(sv);
            const double x = t1 * sv[0];
#pragma unroll
            for (int i = 0; i < EPT; i++) b[i] = (tid + i * NT < K) ? cg * (x + t0 * b[i]) : 0.0;
            Col<NT, EPT>::store(b, beta + (size_t)g * Kp, K);
#pragma unroll
            for (int i = 0; i < EPT; i++) ee[i] = een[i];
        }
    }
}

// gamma -> hapProbs / genProbs.  grid = (T, jobs), 256 threads = 32 SNPs of the grid x 8 slices of K.
// first = this is the first sampling sweep (assign), otherwise accumulate; scale = 1 / n_sample applied on the last.
__global__ void __launch_bounds__(256) k_happrobs(BatchParams P, const JobDev* __restrict__ jobs, int first, int last, double scale) {
    extern __shared__ __align__(16) unsigned char hsm[];
    const JobDev& J = jobs[blockIdx.y];
    if (*J.underflow) return;
    const int g = blockIdx.x, K = P.K, Kp = P.Kp, T = P.T, NH = P.NH, nSNPs = P.nSNPs;
    double* gam = reinterpret_cast<double*>(hsm);                   // [NH][Kp]
    uint32_t* w = reinterpret_cast<uint32_t*>(gam + (size_t)NH * Kp);  // [Kp]
    double* part = reinterpret_cast<double*>(w + Kp);               // [8][3][2][32]
    const int tid = threadIdx.x;
    for (int h = 0; h < NH; h++) {
        const double x = 1 / ld_cg(J.c + h * T + g);
        const double* a = J.alpha + ((size_t)h * T + g) * Kp;
        const double* b = J.beta + ((size_t)h * T + g) * Kp;
        for (int k = tid; k < K; k += 256) gam[h * Kp + k] = (ld_stream(a + k) * ld_stream(b + k)) * x;
    }
    for (int k = tid; k < K; k += 256) w[k] = J.W[(size_t)g * Kp + k];
    __syncthreads();
    const int bit = tid & 31, sl = tid >> 5;
    double alt[3] = {0, 0, 0}, ref[3] = {0, 0, 0};
    const int per = (K + 7) / 8;
    const int k0 = sl * per, k1 = min(K, k0 + per);
    for (int k = k0; k < k1; k++) {
        const bool set = (w[k] >> bit) & 1u;
        for (int h = 0; h < NH; h++) {
            const double gk = gam[h * Kp + k];
            if (set)
                alt[h] += gk;
            else
                ref[h] += gk;
        }
    }
    for (int h = 0; h < 3; h++) {
        part[((sl * 3 + h) * 2 + 0) * 32 + bit] = alt[h];
        part[((sl * 3 + h) * 2 + 1) * 32 + bit] = ref[h];
    }
    __syncthreads();
    if (tid < 32) {
        const int s = 32 * g + bit;
        if (s < nSNPs) {
            double A[3], Rf[3];
            for (int h = 0; h < 3; h++) {
                A[h] = 0;
                Rf[h] = 0;
                for (int q = 0; q < 8; q++) {
                    A[h] += part[((q * 3 + h) * 2 + 0) * 32 + bit];
                    Rf[h] += part[((q * 3 + h) * 2 + 1) * 32 + bit];
                }
            }
            const double eps = P.ref_error, ome = 1 - P.ref_error;
            double hp[3], gM[3], gF[3] = {0, 0, 0};
            if (!P.rare_common) {
                for (int h = 0; h < 3; h++) hp[h] = A[h] * ome + Rf[h] * eps;
                gM[0] = (1 - hp[0]) * (1 - hp[1]);
                gM[1] = (hp[0] * (1 - hp[1]) + (1 - hp[0]) * hp[1]);
                gM[2] = hp[0] * hp[1];
                gF[0] = (1 - hp[0]) * (1 - hp[2]);
                gF[1] = (hp[0] * (1 - hp[2]) + (1 - hp[0]) * hp[2]);
                gF[2] = hp[0] * hp[2];
            } else {
                // the reference accumulates into its (never re-zeroed) local matrix: hapLocal carries that state
                const int type = J.snp_type[s];
                for (int h = 0; h < 3; h++) {
                    double v = J.hapLocal[h * (size_t)nSNPs + s];
                    if (h < NH) {
                        if (type == 0)
                            v += A[h] * ome + Rf[h] * eps;
                        else if (type == 1)
                            v = eps;
                        else
                            v += (A[h] + Rf[h]) * eps + A[h] * (1 - 2 * eps);
                    }
                    hp[h] = v;
                    J.hapLocal[h * (size_t)nSNPs + s] = v;
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
            for (int h = 0; h < 3; h++) {
                // output layout [3 x nSNPs] column-major (row = haplotype / genotype)
                const size_t o = (size_t)s * 3 + h;
                double vh = hp[h], vm = gM[h], vf = gF[h];
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
