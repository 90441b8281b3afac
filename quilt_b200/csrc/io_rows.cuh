// io_rows.cuh — the steps on either side of the Gibbs path.
//
// Ingestion (rank 3): the pileup of one sample -> the reference's sampleReads in path order.
//   QUILT/R/functions.R:243-316: loadBamAndConvert -> sampleReads (list(J, central SNP, bq, u) per read), get_alleleCount
//   (:2779-2800, increment2N QUILT/src/copied-from-stitch.cpp:573-579), snap_sampleReads_to_grid (STITCH, un-vendored: the read's
//   second field becomes grid[central SNP], reads end up ordered by it — gibbs-nipt.cpp:811 relies on non-decreasing wif),
//   grid_has_read (:314-316).  BAM decoding stays on the CPU; what arrives here is the flat pileup.
//   Every sum is taken in the reference's order (increment2N walks the read-SNP entries in read order), so the allele counts are
//   bit-identical: entries are bucketed per SNP / reads per grid with integer atomics, every small bucket is then put back into
//   entry order by one thread, and only then are the floating-point values added.
//
// VCF column (rank 4): per-SNP text of one sample, "GT:GP:DS:HD" = a|b:%.3f,%.3f,%.3f:%.3f:%.3f,%.3f
//   QUILT/R/functions.R:1408-1463: STITCH::rcpp_make_column_of_vcf(gp_t, use_state_probabilities = TRUE, q_t = t(phasing_haps))
//   (un-vendored; format from the QUILT VCF header QUILT/R/writers.R:10-36 and its call site) with the unphased GT replaced by
//   round(phasing_haps[, 1]) | round(phasing_haps[, 2]) (:1436-1442).  printf("%.3f") semantics (round half to even on the EXACT
//   value) are reproduced with an fma residual, not with rint(1000 x).
#pragma once

#include "device_common.cuh"
#include "types.h"

namespace qb {

struct IngestDev {
    int32_t R, nU, nSNPs, T;
    const int32_t* off;      // [R + 1] pileup order
    const int32_t* u;        // [nU] 0-based SNP
    const int32_t* bq;       // [nU] signed
    const int32_t* central;  // [R] 0-based central SNP of the read (sampleReads[[r]][[2]] before snapping)
    const int32_t* grid;     // [nSNPs] 0-based grid of every SNP
    const double* prob;      // [nU][2] convertScaledBQtoProbs of bq (host libm pow)
    int32_t* wif;            // [R] scratch: grid of the central SNP
    int32_t* cntG;           // [T + 1] reads per grid -> first read per grid (rs)
    int32_t* fillG;          // [T]
    int32_t* ordG;           // [R] reads in path order
    int32_t* cntS;           // [nSNPs + 1] entries per SNP -> first entry per SNP
    int32_t* fillS;          // [nSNPs]
    int32_t* ordS;           // [nU] entries bucketed per SNP
    int32_t* ncnt;           // [R + 1] SNPs per read in path order -> offsets_sorted
    // outputs
    int32_t* u_s;            // [nU]
    int32_t* bq_s;           // [nU]
    int32_t* wif_s;          // [R]
    double* alleleCount;     // [nSNPs][2] column-major: (alt count, total count)
    uint8_t* grid_has_read;  // [T]
};

__global__ void __launch_bounds__(256) k_ing_count(IngestDev D) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < D.R) {
        const int w = D.grid[D.central[i]];
        D.wif[i] = w;
        atomicAdd(&D.cntG[w], 1);
    }
    if (i < D.nU) atomicAdd(&D.cntS[D.u[i]], 1);
}

// in-place exclusive prefix sum of a[0 .. n) with the total in a[n]; one CTA of 1024 threads
__global__ void __launch_bounds__(1024) k_exscan_i32(int32_t* __restrict__ a, int n) {
    __shared__ int wsum[32];
    __shared__ int carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + tid;
        const int v = (i < n) ? a[i] : 0;
        int incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += y;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        int wo = 0;
        for (int w = 0; w < warp; w++) wo += wsum[w];
        const int c = carry;
        if (i < n) a[i] = c + wo + incl - v;
        __syncthreads();
        if (tid == 1023) carry = c + wo + incl;
        __syncthreads();
    }
    if (tid == 0) a[n] = carry;
}

__global__ void __launch_bounds__(256) k_ing_place(IngestDev D) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < D.R) {
        const int w = D.wif[i];
        D.ordG[D.cntG[w] + atomicAdd(&D.fillG[w], 1)] = i;
    }
    if (i < D.nU) {
        const int s = D.u[i];
        D.ordS[D.cntS[s] + atomicAdd(&D.fillS[s], 1)] = i;
    }
}

// every bucket back into entry order (insertion sort: buckets hold a grid's reads / a SNP's entries, tens of items)
__global__ void __launch_bounds__(256) k_ing_sort(IngestDev D) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    for (int pass = 0; pass < 2; pass++) {
        const int n = pass == 0 ? D.T : D.nSNPs;
        if (i >= n) continue;
        const int32_t* first = pass == 0 ? D.cntG : D.cntS;
        int32_t* ord = pass == 0 ? D.ordG : D.ordS;
        const int a = first[i], b = first[i + 1];
        for (int x = a + 1; x < b; x++) {
            const int v = ord[x];
            int y = x - 1;
            while (y >= a && ord[y] > v) {
                ord[y + 1] = ord[y];
                y--;
            }
            ord[y + 1] = v;
        }
    }
    if (i < D.T) D.grid_has_read[i] = D.cntG[i + 1] > D.cntG[i] ? 1 : 0;
}

__global__ void __launch_bounds__(256) k_ing_lens(IngestDev D) {
    const int q = blockIdx.x * 256 + threadIdx.x;
    if (q < D.R) {
        const int r = D.ordG[q];
        D.ncnt[q] = D.off[r + 1] - D.off[r];
        D.wif_s[q] = D.wif[r];
    }
}

__global__ void __launch_bounds__(256) k_ing_copy(IngestDev D) {
    const int q = blockIdx.x;  // one CTA per read in path order
    const int r = D.ordG[q];
    const int a = D.off[r], n = D.off[r + 1] - a, o = D.ncnt[q];
    for (int j = threadIdx.x; j < n; j += 256) {
        D.u_s[o + j] = D.u[a + j];
        D.bq_s[o + j] = D.bq[a + j];
    }
}

// get_alleleCount: c1[s] = sum of prob[t][0], c2[s] = sum of prob[t][1] over the entries t of SNP s in entry order;
// alleleCount[, 1] = c2, alleleCount[, 2] = c1 + c2
__global__ void __launch_bounds__(256) k_ing_allele(IngestDev D) {
    const int s = blockIdx.x * 256 + threadIdx.x;
    if (s >= D.nSNPs) return;
    double c1 = 0, c2 = 0;
    for (int x = D.cntS[s]; x < D.cntS[s + 1]; x++) {
        const int t = D.ordS[x];
        c1 = c1 + D.prob[2 * (size_t)t];
        c2 = c2 + D.prob[2 * (size_t)t + 1];
    }
    D.alleleCount[s] = c2;
    D.alleleCount[(size_t)D.nSNPs + s] = c1 + c2;
}

// ------------------------------------------------------------------------------------------------ VCF column
constexpr int VCF_REC = 39;  // a|b:0.000,0.000,0.000:0.000:0.000,0.000

// printf("%.3f", x) for 0 <= x < 10: thousandths of x rounded half-to-even on the exact value
__device__ __forceinline__ int thousandths_exact(double x) {
    const double p = x * 1000.0;
    const double err = fma(x, 1000.0, -p);  // x * 1000 = p + err exactly
    const double n = floor(p);
    const double frac = p - n;  // exact
    int k = (int)n;
    // (p is the nearest double to the exact product and n + 0.5 is a double: frac is on the same side of 0.5 as the exact
    //  value, except exactly at 0.5, where the residual decides; an exact tie goes to the even neighbour like glibc's printf)
    if (frac > 0.5 || (frac == 0.5 && (err > 0 || (err == 0 && (k & 1))))) k++;
    return k;
}
__device__ __forceinline__ void put_fixed3(char* o, double x) {
    int k = thousandths_exact(x);
    if (k > 9999) k = 9999;
    if (k < 0) k = 0;
    o[0] = (char)('0' + k / 1000);
    o[1] = '.';
    o[2] = (char)('0' + (k / 100) % 10);
    o[3] = (char)('0' + (k / 10) % 10);
    o[4] = (char)('0' + k % 10);
}

// gp [3][nSNPs] column-major (element (i, s) at 3 s + i), hd [nSNPs x 2] column-major; out [nSNPs][VCF_REC] bytes (no terminator)
__global__ void __launch_bounds__(256) k_vcf_column(int nSNPs, const double* __restrict__ gp, const double* __restrict__ hd, char* __restrict__ out) {
    const int s = blockIdx.x * 256 + threadIdx.x;
    if (s >= nSNPs) return;
    const double g0 = gp[3 * (size_t)s], g1 = gp[3 * (size_t)s + 1], g2 = gp[3 * (size_t)s + 2];
    const double h1 = hd[s], h2 = hd[(size_t)nSNPs + s];
    char r[VCF_REC];
    r[0] = (char)('0' + (int)rint(h1));
    r[1] = '|';
    r[2] = (char)('0' + (int)rint(h2));
    r[3] = ':';
    put_fixed3(r + 4, g0);
    r[9] = ',';
    put_fixed3(r + 10, g1);
    r[15] = ',';
    put_fixed3(r + 16, g2);
    r[21] = ':';
    put_fixed3(r + 22, g1 + 2 * g2);  // dosage
    r[27] = ':';
    put_fixed3(r + 28, h1);
    r[33] = ',';
    put_fixed3(r + 34, h2);
    char* o = out + (size_t)s * VCF_REC;
#pragma unroll
    for (int i = 0; i < VCF_REC; i++) o[i] = r[i];
}

}  // namespace qb
