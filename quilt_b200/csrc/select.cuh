// select.cuh — haplotype re-selection between Gibbs calls on the device.
//
// Reference: select_new_haps_mspbwt_v3, heuristic_approach "A" (QUILT/R/mspbwt.R:230-474), called between the
// common-SNP Gibbs calls of a chain (QUILT/R/functions.R:856-868).  The in-tree R (rounding and packing :277-278, the
// interleaved grid subsets :283, the per-subset ordering / de-duplication :312-335, the -len1 ordering :349, the
// coverage-weighted ranking :414-441, the interleave / unique / first-Knew cut :443-466) is reproduced exactly; the two
// calls into the un-vendored mspbwt package (map_Z_to_all_symbols, Rcpp_find_good_matches_without_a) follow the contract
// written down in include/quilt_b200.h (QuiltSelectArgs) — parity unpinned for that part.
//
//  k_sel_symbols   hapProbs -> rounded 32-SNP words -> the panel's symbol of each word          grid (Tc, nHap, jobs)
//  k_sel_match     per (subset, haplotype of the sample): run lengths of all K_full panel haplotypes, the 2L longest per
//                  position, rows (index1, start1, end1) of the runs that end as neighbours     grid (nIndices, nHap, jobs)
//  k_sel_rank      everything after the matching, one CTA per job: sorts, de-duplication, first occurrences, weights,
//                  interleave, cut, optional completion of a short list                         grid (jobs)
#pragma once

#include "device_common.cuh"
#include "types.h"

namespace qb {

constexpr int SEL_NT = 256;        // threads of k_sel_symbols / k_sel_match
constexpr int SEL_RANK_NT = 1024;  // threads of k_sel_rank
constexpr int SEL_MAXTOP = 16;     // 2 * mspbwtL <= 16

struct SelParams {
    int32_t K_full, Tc, nSNPs, nMaxDH;
    int32_t nHap, nIndices, L, M, Knew;
    int32_t rows_cap;    // capacity of one (haplotype, subset) row list = positions of the longest subset * 2L
    int32_t sort_cap;    // power of two >= nIndices * rows_cap: capacity of the per-haplotype sort buffers
    int32_t pad;         // 1: complete a short list from the haplotypes not in it (chain mode)
};

// per job: where its inputs / outputs / scratch live (device pointers)
struct SelJob {
    const double* hapProbs;  // [3][nSNPs] column-major (element (h, s) at 3 s + h)
    int32_t* which_out;      // [Knew] 1-based result
    int32_t* counts_out;     // [2] n_found, n_unique
    const double* pad_unif;  // [Knew] (chain mode) or null
    // scratch
    int32_t* sym;            // [nHap][Tc]
    uint16_t* runlen;        // [nHap][nIndices][K_full]
    uint64_t* rows;          // [nHap][nIndices][rows_cap]   packed (index1 << 40 | start1 << 20 | end1)
    int32_t* nrows;          // [nHap][nIndices]
    uint64_t* skey;          // [sort_cap]
    uint32_t* sval;          // [sort_cap]
    uint64_t* list;          // [nHap][sort_cap] rows of each haplotype after the -len1 ordering
    int32_t* nlist;          // [nHap]
    int32_t* firstpos;       // [max(K_full, Tc) + 2]
    int32_t* tmp;            // [max(nHap * sort_cap, K_full + 2)] flags / prefix sums
    int32_t* ranked;         // [nHap * sort_cap] haplotype sequences (concatenated / interleaved)
    int32_t* tmp2;           // [nHap * sort_cap] per-haplotype ranked lists (weighted branch)
    uint64_t* stage;         // [sort_cap] gather staging
    int32_t* pool;           // [K_full + 2] free haplotypes (chain-mode completion)
};

__device__ __forceinline__ uint64_t sel_pack(int index1, int start1, int end1) { return ((uint64_t)index1 << 40) | ((uint64_t)start1 << 20) | (uint64_t)end1; }
__device__ __forceinline__ int sel_index1(uint64_t r) { return (int)(r >> 40); }
__device__ __forceinline__ int sel_start1(uint64_t r) { return (int)((r >> 20) & 0xfffff); }
__device__ __forceinline__ int sel_end1(uint64_t r) { return (int)(r & 0xfffff); }

// round(hapProbs_t[h, ]) packed 32 SNPs per word (R's round is round-half-even, rint() here), then the row of
// distinctHapsB[, g] holding that word among the rows in use (-1: the word is not in the panel's table)
__global__ void __launch_bounds__(SEL_NT) k_sel_symbols(SelParams P, const SelJob* __restrict__ jobs, const int32_t* __restrict__ distinctHapsB,
                                                        const int32_t* __restrict__ n_used) {
    const SelJob& J = jobs[blockIdx.z];
    const int g = blockIdx.x, h = blockIdx.y, tid = threadIdx.x;
    __shared__ uint32_t word;
    __shared__ int best;
    if (tid < 32) {
        const int s = 32 * g + tid;
        const bool bit = (s < P.nSNPs) && (rint(J.hapProbs[(size_t)s * 3 + h]) != 0.0);
        const uint32_t w = __ballot_sync(0xffffffffu, bit);
        if (tid == 0) {
            word = w;
            best = 0x7fffffff;
        }
    }
    __syncthreads();
    const int nu = n_used[g];
    for (int i = tid; i < nu; i += SEL_NT)
        if ((uint32_t)distinctHapsB[(size_t)g * P.nMaxDH + i] == word) atomicMin(&best, i);
    __syncthreads();
    if (tid == 0) J.sym[h * P.Tc + g] = (best == 0x7fffffff) ? -1 : best + 1;
}

// one CTA walks one subset of grids for one haplotype of the sample
__global__ void __launch_bounds__(SEL_NT) k_sel_match(SelParams P, const SelJob* __restrict__ jobs, const uint8_t* __restrict__ hapMatcherR) {
    const SelJob& J = jobs[blockIdx.z];
    const int iIndex = blockIdx.x, h = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = P.K_full, nI = P.nIndices, top_n = 2 * P.L;
    const int n_pos = (P.Tc - iIndex + nI - 1) / nI;  // length of seq(iIndex + 1, nGrids, nIndices)
    uint16_t* len = J.runlen + ((size_t)h * nI + iIndex) * K;
    uint64_t* rows = J.rows + ((size_t)h * nI + iIndex) * P.rows_cap;
    const int32_t* sym = J.sym + h * P.Tc;
    __shared__ unsigned long long wtop[SEL_NT / 32][SEL_MAXTOP];
    __shared__ unsigned long long top[SEL_MAXTOP];
    __shared__ int n_out;
    for (int k = tid; k < K; k += SEL_NT) len[k] = 0;
    if (tid == 0) n_out = 0;
    __syncthreads();
    for (int j = 0; j < n_pos; j++) {
        const int g = iIndex + j * nI;
        const uint8_t* col = hapMatcherR + (size_t)g * K;
        const int zs = sym[g];
        // 1. run lengths, local top (sorted descending); key = run length, then lower haplotype index
        unsigned long long loc[SEL_MAXTOP];
#pragma unroll
        for (int t = 0; t < SEL_MAXTOP; t++) loc[t] = 0ull;
        for (int k = tid; k < K; k += SEL_NT) {
            const int l = (zs > 0 && col[k] == zs) ? len[k] + 1 : 0;
            len[k] = (uint16_t)l;
            if (l > 0) {
                unsigned long long key = ((unsigned long long)l << 32) | (unsigned)(0x7fffffff - k);
                if (key > loc[SEL_MAXTOP - 1]) {
#pragma unroll
                    for (int t = 0; t < SEL_MAXTOP; t++) {
                        if (key > loc[t]) {
                            const unsigned long long o = loc[t];
                            loc[t] = key;
                            key = o;
                        }
                    }
                }
            }
        }
        // 2. warp top: repeated shuffle-max, the owner pops its head
        int head = 0;
        for (int r = 0; r < top_n; r++) {
            unsigned long long cand = 0ull;
#pragma unroll
            for (int t = 0; t < SEL_MAXTOP; t++)
                if (t == head) cand = loc[t];
            unsigned long long m = cand;
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) {
                const unsigned long long o = __shfl_xor_sync(0xffffffffu, m, d);
                m = o > m ? o : m;
            }
            if (m != 0ull && cand == m) head++;  // keys are unique (they embed the haplotype index)
            if (lane == 0) wtop[warp][r] = m;
        }
        __syncthreads();
        // 3. block top: warp 0 merges the (SEL_NT / 32) * top_n warp keys the same way
        if (warp == 0) {
            constexpr int NW = SEL_NT / 32;
            unsigned long long mine[(NW * SEL_MAXTOP + 31) / 32];
            const int total = NW * top_n;
#pragma unroll
            for (int q = 0; q < (NW * SEL_MAXTOP + 31) / 32; q++) {
                const int e = lane + 32 * q;
                mine[q] = (e < total) ? wtop[e / top_n][e % top_n] : 0ull;
            }
            for (int r = 0; r < top_n; r++) {
                unsigned long long cand = 0ull;
#pragma unroll
                for (int q = 0; q < (NW * SEL_MAXTOP + 31) / 32; q++) cand = mine[q] > cand ? mine[q] : cand;
                unsigned long long m = cand;
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) {
                    const unsigned long long o = __shfl_xor_sync(0xffffffffu, m, d);
                    m = o > m ? o : m;
                }
                if (m != 0ull) {
#pragma unroll
                    for (int q = 0; q < (NW * SEL_MAXTOP + 31) / 32; q++)
                        if (mine[q] == m) mine[q] = 0ull;
                }
                if (lane == 0) top[r] = m;
            }
        }
        __syncthreads();
        // 4. neighbours whose run ends here are reported, in neighbour order (deterministic)
        if (tid == 0) {
            const bool last = (j == n_pos - 1);
            const uint8_t* ncol = last ? nullptr : hapMatcherR + (size_t)(g + nI) * K;
            const int nzs = last ? -1 : sym[g + nI];
            for (int r = 0; r < top_n; r++) {
                const unsigned long long m = top[r];
                if (m == 0ull) break;
                const int k = 0x7fffffff - (int)(m & 0xffffffffull), l = (int)(m >> 32);
                const bool ends = last || !(nzs > 0 && ncol[k] == nzs);
                if (ends && l >= P.M && n_out < P.rows_cap) rows[n_out++] = sel_pack(k + 1, (j - l + 1) + 1, (j - l + 1) + l);
            }
        }
        __syncthreads();
    }
    if (tid == 0) J.nrows[h * nI + iIndex] = n_out;
}

// ---- single-CTA helpers (all threads of the CTA call them; arrays live in global memory, sizes are small)
// ascending sort of (key, val) pairs, n2 a power of two
__device__ inline void cta_bitonic(uint64_t* key, uint32_t* val, int n2) {
    for (int k = 2; k <= n2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const bool up = (i & k) == 0;
                    const uint64_t a = key[i], b = key[ixj];
                    const uint32_t va = val[i], vb = val[ixj];
                    const bool gt = a > b || (a == b && va > vb);
                    if (gt == up) {
                        key[i] = b;
                        key[ixj] = a;
                        val[i] = vb;
                        val[ixj] = va;
                    }
                }
            }
            __syncthreads();
        }
    }
}
// flag[i] (0 / 1) -> exclusive prefix sum in place; returns the total.  scr: shared int[blockDim.x + 1]
__device__ inline int cta_scan_flags(int32_t* flag, int n, int* scr) {
    const int nt = blockDim.x, tid = threadIdx.x;
    const int per = (n + nt - 1) / nt;
    const int a = min(tid * per, n), b = min(a + per, n);
    int c = 0;
    for (int i = a; i < b; i++) c += flag[i];
    scr[tid + 1] = c;
    if (tid == 0) scr[0] = 0;
    __syncthreads();
    if (tid < 32) {
        // warp 0 scans the per-thread counts in chunks of 32
        int carry = 0;
        for (int base = 0; base < nt; base += 32) {
            int v = scr[base + tid + 1];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(0xffffffffu, v, d);
                if (tid >= d) v += o;
            }
            scr[base + tid + 1] = v + carry;
            carry += __shfl_sync(0xffffffffu, v, 31);
        }
    }
    __syncthreads();
    int run = scr[tid];
    for (int i = a; i < b; i++) {
        const int f = flag[i];
        flag[i] = run;
        run += f;
    }
    const int total = scr[nt];
    __syncthreads();
    return total;
}
__device__ inline int next_pow2(int n) {
    int p = 1;
    while (p < n) p <<= 1;
    return p;
}

// unique() with first occurrences kept over seq[0..n) (entries < 0 are NA and dropped): out[0..ret) in order of first
// appearance.  firstpos: int[K_full + 1] scratch, flag: int[n] scratch
__device__ inline int cta_unique_keep_first(const int32_t* seq, int n, int K_full, int32_t* firstpos, int32_t* flag, int32_t* out, int out_cap, int* scr) {
    for (int k = threadIdx.x; k <= K_full; k += blockDim.x) firstpos[k] = 0x7fffffff;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        if (seq[i] > 0) atomicMin(&firstpos[seq[i]], i);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) flag[i] = (seq[i] > 0 && firstpos[seq[i]] == i) ? 1 : 0;
    __syncthreads();
    // (flags are overwritten by their prefix sums; keep the decision in the sign of firstpos: firstpos == i <=> kept)
    const int total = cta_scan_flags(flag, n, scr);
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        if (seq[i] > 0 && firstpos[seq[i]] == i && flag[i] < out_cap) out[flag[i]] = seq[i];
    __syncthreads();
    return total;
}

__global__ void __launch_bounds__(SEL_RANK_NT) k_sel_rank(SelParams P, const SelJob* __restrict__ jobs) {
    const SelJob& J = jobs[blockIdx.x];
    const int tid = threadIdx.x, nI = P.nIndices, nH = P.nHap;
    __shared__ int scr[SEL_RANK_NT + 1];
    __shared__ int s_n;
    // ---- per haplotype: subsets ordered / de-duplicated (mspbwt.R:323-335), concatenated, ordered by -len1 (:349)
    for (int h = 0; h < nH; h++) {
        int n_list = 0;
        uint64_t* list = J.list + (size_t)h * P.sort_cap;
        for (int i = 0; i < nI; i++) {
            const int n = J.nrows[h * nI + i];
            const uint64_t* rows = J.rows + ((size_t)h * nI + i) * P.rows_cap;
            if (n == 0) continue;
            const int n2 = next_pow2(n);
            // order(index1, -end1, -start1): one composite ascending key (rows of a subset are distinct in it)
            for (int q = tid; q < n2; q += SEL_RANK_NT) {
                uint64_t key = ~0ull;
                if (q < n) {
                    const uint64_t r = rows[q];
                    key = ((uint64_t)sel_index1(r) << 40) | ((uint64_t)(0xfffff - sel_end1(r)) << 20) | (uint64_t)(0xfffff - sel_start1(r));
                }
                J.skey[q] = key;
                J.sval[q] = (uint32_t)q;
            }
            __syncthreads();
            cta_bitonic(J.skey, J.sval, n2);
            // x <- c(FALSE, diff(index1) == 0 & diff(start1) == 0); keep !x
            for (int q = tid; q < n; q += SEL_RANK_NT) {
                int keep = 1;
                if (q > 0) {
                    const uint64_t a = J.skey[q], b = J.skey[q - 1];
                    if ((a >> 40) == (b >> 40) && (a & 0xfffff) == (b & 0xfffff)) keep = 0;
                }
                J.tmp[q] = keep;
            }
            __syncthreads();
            // (the decision must survive the in-place scan: recompute it when scattering)
            const int kept = cta_scan_flags(J.tmp, n, scr);
            for (int q = tid; q < n; q += SEL_RANK_NT) {
                bool keep = true;
                if (q > 0) {
                    const uint64_t a = J.skey[q], b = J.skey[q - 1];
                    keep = !((a >> 40) == (b >> 40) && (a & 0xfffff) == (b & 0xfffff));
                }
                if (keep) {
                    const uint64_t a = J.skey[q];
                    list[n_list + J.tmp[q]] = sel_pack((int)(a >> 40), 0xfffff - (int)(a & 0xfffff), 0xfffff - (int)((a >> 20) & 0xfffff));
                }
            }
            __syncthreads();
            n_list += kept;
        }
        // mtm[order(-len1), ] (stable: ties keep the concatenation order)
        const int n2 = next_pow2(max(n_list, 1));
        for (int q = tid; q < n2; q += SEL_RANK_NT) {
            uint64_t key = ~0ull;
            if (q < n_list) {
                const uint64_t r = list[q];
                key = (uint64_t)(0xfffff - (sel_end1(r) - sel_start1(r) + 1));
            }
            J.skey[q] = key;
            J.sval[q] = (uint32_t)q;
        }
        __syncthreads();
        cta_bitonic(J.skey, J.sval, n2);
        // gather through the permutation
        uint64_t* stage = J.stage;
        for (int q = tid; q < n_list; q += SEL_RANK_NT) stage[q] = list[J.sval[q]];
        __syncthreads();
        for (int q = tid; q < n_list; q += SEL_RANK_NT) list[q] = stage[q];
        __syncthreads();
        if (tid == 0) J.nlist[h] = n_list;
        __syncthreads();
    }
    // ---- unique_haps <- unique(c(out[[1]][, 1], out[[2]][, 1] (, out[[3]][, 1])))   (:352-356)
    int n_cat = 0;
    for (int h = 0; h < nH; h++) {
        const int n = J.nlist[h];
        const uint64_t* list = J.list + (size_t)h * P.sort_cap;
        for (int q = tid; q < n; q += SEL_RANK_NT) J.ranked[n_cat + q] = sel_index1(list[q]);
        n_cat += n;
    }
    __syncthreads();
    // (seq = ranked[0 .. n_cat), flags in tmp, result straight into which_out)
    const int n_unique = cta_unique_keep_first(J.ranked, n_cat, P.K_full, J.firstpos, J.tmp, J.which_out, P.Knew, scr);
    int n_found = min(n_unique, P.Knew);
    if (n_unique > P.Knew) {
        // ---- coverage-weighted ranking per haplotype (:414-441), then interleave / unique / first Knew (:443-466)
        int amax = 0;
        for (int h = 0; h < nH; h++) {
            const int n = J.nlist[h];
            uint64_t* list = J.list + (size_t)h * P.sort_cap;
            amax = max(amax, n);
            // cur_sum over 1 .. max(end1): small integers, kept as int32 in firstpos[] (sums are exact in any order)
            int32_t* cur = J.firstpos;
            for (int q = tid; q <= P.Tc; q += SEL_RANK_NT) cur[q] = 1;
            __syncthreads();
            if (tid < 32) {
                for (int i = 0; i < n; i++) {
                    const uint64_t r = list[i];
                    const int s = sel_start1(r), e = sel_end1(r);
                    int acc = 0;
                    for (int q = s + tid; q <= e; q += 32) acc += cur[q];
#pragma unroll
                    for (int d = 16; d >= 1; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
                    __syncwarp();
                    for (int q = s + tid; q <= e; q += 32) cur[q] += 1;
                    __syncwarp();
                    if (tid == 0) {
                        const double w = ((double)(e - s + 1) * 1) / (double)acc;  // (e - s + 1) * 1 / sum(cur_sum[s:e])
                        // order(-weight), stable: positive doubles order like their bit patterns
                        J.skey[i] = ~(uint64_t)__double_as_longlong(w);
                        J.sval[i] = (uint32_t)i;
                    }
                }
            }
            __syncthreads();
            const int n2 = next_pow2(max(n, 1));
            for (int q = n + tid; q < n2; q += SEL_RANK_NT) {
                J.skey[q] = ~0ull;
                J.sval[q] = 0xffffffffu;
            }
            __syncthreads();
            cta_bitonic(J.skey, J.sval, n2);
            for (int q = tid; q < n; q += SEL_RANK_NT) J.tmp2[(size_t)h * P.sort_cap + q] = sel_index1(list[J.sval[q]]);
            __syncthreads();
        }
        // c(t(cbind(x, y (, z)))) with NA padding
        for (int q = tid; q < amax * nH; q += SEL_RANK_NT) {
            const int i = q / nH, h = q - i * nH;
            J.ranked[q] = (i < J.nlist[h]) ? J.tmp2[(size_t)h * P.sort_cap + i] : -1;
        }
        __syncthreads();
        cta_unique_keep_first(J.ranked, amax * nH, P.K_full, J.firstpos, J.tmp, J.which_out, P.Knew, scr);
        n_found = P.Knew;
    }
    // ---- chain mode: complete a short list from the haplotypes not in it (increasing order) by a partial Fisher-Yates
    if (P.pad && n_found < P.Knew) {
        for (int k = tid; k <= P.K_full; k += SEL_RANK_NT) J.firstpos[k] = 1;  // 1 = still free
        __syncthreads();
        for (int q = tid; q < n_found; q += SEL_RANK_NT) J.firstpos[J.which_out[q]] = 0;
        if (tid == 0) J.firstpos[0] = 0;
        __syncthreads();
        // pool = free haplotypes in increasing order: flags -> positions (tmp), scatter into ranked[]
        for (int k = tid; k <= P.K_full; k += SEL_RANK_NT) J.tmp[k] = J.firstpos[k];
        __syncthreads();
        const int n_pool = cta_scan_flags(J.tmp, P.K_full + 1, scr);
        for (int k = tid; k <= P.K_full; k += SEL_RANK_NT)
            if (J.firstpos[k]) J.pool[J.tmp[k]] = k;
        __syncthreads();
        if (tid == 0) {
            int n_left = n_pool, t = 0, nf = n_found;
            while (nf < P.Knew && n_left > 0) {
                int j = (int)floor(n_left * J.pad_unif[t++]);
                if (j >= n_left) j = n_left - 1;
                J.which_out[nf++] = J.pool[j];
                J.pool[j] = J.pool[--n_left];
            }
            s_n = nf;
        }
        __syncthreads();
        n_found = s_n;
    }
    for (int q = n_found + tid; q < P.Knew; q += SEL_RANK_NT) J.which_out[q] = 0;
    if (tid == 0) {
        J.counts_out[0] = n_found;
        J.counts_out[1] = n_unique;
    }
}

// ------------------------------------------------------------------------------------------------ per-sample summary
// R-side reductions after the Gibbs calls of a sample (QUILT/R/functions.R:999-1020, :1304-1325, recast_haps :3180-3209,
// eij / fij / max_gen :1408-1411), one thread per SNP, the stored calls added in the reference's order.
struct SumSample {
    const double* const* call_hp;  // [n_calls] device pointers to hapProbs_t ([3][nSNPs] column-major)
    const double* phase_hp;
    int32_t n_calls;
    double* dosage;   // [nSNPs]
    double* gp;       // [3][nSNPs] column-major (element (i, s) at 3 s + i)
    double* hd;       // [2][nSNPs] -> column-major [nSNPs x 2]
    int8_t* gt;       // [nSNPs x 2]
    double* eij;      // [nSNPs]
    double* fij;      // [nSNPs]
    int8_t* maxgen;   // [nSNPs] 0-based arg-max genotype
};
__device__ __forceinline__ double r_round3(double x) { return rint(x * 1000.0) / 1000.0; }

// grid = (ceil(nSNPs / 256), samples)
__global__ void __launch_bounds__(256) k_sample_summary(const SumSample* __restrict__ S, int nSNPs) {
    const SumSample& J = S[blockIdx.y];
    const int s = blockIdx.x * 256 + threadIdx.x;
    if (s >= nSNPs) return;
    double dosage = 0, g0 = 0, g1 = 0, g2 = 0;
    for (int c = 0; c < J.n_calls; c++) {
        const double* hp = J.call_hp[c];
        const double hap1 = hp[3 * (size_t)s], hap2 = hp[3 * (size_t)s + 1];
        dosage = dosage + hap1 + hap2;                               // dosage <- dosage + hap1 + hap2
        g0 = g0 + (1 - hap1) * (1 - hap2);                           // gp_t <- gp_t + rbind(...)
        g1 = g1 + ((1 - hap1) * hap2 + hap1 * (1 - hap2));
        g2 = g2 + hap1 * hap2;
    }
    // recast_haps(hd1, hd2, gp = t(gp_t)) with the sums as they stand at the phasing iteration (the arg-max is scale-free)
    double hd1 = J.phase_hp[3 * (size_t)s], hd2 = J.phase_hp[3 * (size_t)s + 1];
    {
        const double gt1 = rint(hd1) + rint(hd2);
        double max_val = g0;
        int gt3 = 0;
        if (g1 > max_val) {
            gt3 = 1;
            max_val = g1;
        }
        if (g2 > max_val) {
            gt3 = 2;
            max_val = g2;
        }
        if ((double)gt3 != gt1) {
            if (gt3 == 0) {
                hd1 = 0;
                hd2 = 0;
            } else if (gt3 == 2) {
                hd1 = 1;
                hd2 = 1;
            } else {
                const bool first = hd1 > hd2;
                hd1 = first ? 1 : 0;
                hd2 = first ? 0 : 1;
            }
        }
    }
    const double n = (double)J.n_calls;
    dosage = dosage / n;
    g0 = g0 / n;
    g1 = g1 / n;
    g2 = g2 / n;
    J.dosage[s] = dosage;
    J.gp[3 * (size_t)s] = g0;
    J.gp[3 * (size_t)s + 1] = g1;
    J.gp[3 * (size_t)s + 2] = g2;
    J.hd[s] = hd1;
    J.hd[(size_t)nSNPs + s] = hd2;
    J.gt[s] = (int8_t)rint(hd1);
    J.gt[(size_t)nSNPs + s] = (int8_t)rint(hd2);
    J.eij[s] = r_round3(g1 + 2 * g2);
    J.fij[s] = r_round3(g1 + 4 * g2);
    int mg = 0;  // get_max_gen_rapid: the first maximum
    double mv = g0;
    if (g1 > mv) {
        mg = 1;
        mv = g1;
    }
    if (g2 > mv) mg = 2;
    J.maxgen[s] = (int8_t)mg;
}

// per-rank counters over the samples IN ORDER (quilt.R:957-961): one thread per SNP
__global__ void __launch_bounds__(256) k_info_counts(const SumSample* __restrict__ S, int n_samples, int nSNPs, double* __restrict__ info /*[2][nSNPs]*/,
                                                     double* __restrict__ af, double* __restrict__ hwe /*[3][nSNPs]*/) {
    const int s = blockIdx.x * 256 + threadIdx.x;
    if (s >= nSNPs) return;
    double i1 = 0, i2 = 0, a = 0, h[3] = {0, 0, 0};
    for (int q = 0; q < n_samples; q++) {
        const double e = S[q].eij[s], f = S[q].fij[s];
        i1 = i1 + e;
        i2 = i2 + (f - e * e);
        a = a + e / 2;
        const int mg = S[q].maxgen[s];
        h[0] += mg == 0;
        h[1] += mg == 1;
        h[2] += mg == 2;
    }
    info[s] = i1;
    info[(size_t)nSNPs + s] = i2;
    af[s] = a;
    for (int q = 0; q < 3; q++) hwe[(size_t)q * nSNPs + s] = h[q];
}

}  // namespace qb
