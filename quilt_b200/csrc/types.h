// types.h — host/device structures of the B200 Gibbs path (internal to libquiltgpu.so).
#pragma once
#include <stdint.h>

namespace qb {

// A read's emission column E[k, r] (reference: eMatRead_t, gibbs-small.cpp:116-265) takes one of at
// most 2^nb values, selected by the alleles of haplotype k at the read's nb SNPs.  For nb <= NBMAX
// (and SNPs inside grids wif0-1 .. wif0+1) we keep the 2^nb-entry table instead of the K-long fp64
// column (DESIGN.md "emission tables"); everything else falls back to a dense K-long column.
constexpr int NBMAX = 9;
constexpr double ONE_THRESH = 1.0 - 1e-12;  // rcpp_evaluate_read_variability, gibbs-nipt.cpp:349

struct TabEnt {
    double E;     // rescaled, floored emission value for this allele pattern
    double invE;  // 1 / E (used only inside K-long sums; stored state is updated with true divisions)
};

enum ReadMode : uint8_t { MODE_RUN = 0, MODE_GATHER = 1, MODE_DENSE = 2 };

struct ReadDesc {        // 32 bytes, 16-byte aligned (staged to shared memory with cp.async)
    uint32_t off;        // table modes: first TabEnt in the job's table pool; dense: column index in the dense pool
    uint8_t cat;         // read_category after the force_reset / disable flags (gibbs-nipt.cpp:338-382, :2816-2827)
    uint8_t mode;        // ReadMode
    uint8_t nb;          // pattern bits (SNPs used = min(J + 1, Jmax + 1))
    int8_t g0rel;        // MODE_RUN: word index of the first SNP minus wif0 (-1, 0, +1)
    uint8_t b0;          // MODE_RUN: bit of the first SNP inside that word
    uint8_t sel[8];      // MODE_GATHER (no longer produced by the host): per SNP ((word - wif0 + 1) << 5) | bit   (byte offset 9)
    uint8_t pad[3];
    uint32_t tnext;      // table-pool offset just after this read's table (cumulative; unchanged by dense reads)
    // staging chunks of the sweep kernel (host-computed greedy packing of a grid's reads into SW_MAXR reads /
    // SW_MAXTAB table entries): set on the first read of every chunk
    uint16_t chunk_n;    // reads in the chunk that starts at this read (0 elsewhere)
    uint16_t pad2;
    uint32_t chunk_tend; // table-pool offset just after the chunk's last table
};
static_assert(sizeof(ReadDesc) == 32, "ReadDesc must be 32 bytes");

// haplotype classes (k_build_classes / k_sweep): at most CLS_MAX distinct classes per grid, CLS_LANES classes per lane
constexpr int CLS_LANES = 8;
constexpr int CLS_MAX = CLS_LANES * 32 - 1;  // one id is left for the padding elements k >= K
constexpr int CLS_AS = CLS_LANES * 32 + 8;   // stride of the class-sum exchange buffer in shared memory (index = class + 1)

constexpr int LIK_N = 16;  // per-sweep bookkeeping record written by the sweep epilogue
// lik[0..2] = -sum_g log c_h[g] ; lik[3..5] = #reads with label h ; lik[6..13] = #reads with H_class 0..7 ;
// lik[14] = 1 if a non-finite sum(c_h) was seen (reference underflow check, gibbs-nipt.cpp:2959-2969)
// slot 15 of row 0: position in the episode stream (QuiltGibbsArgs.unif_stream) of the three-haplotype kernels
constexpr int LIK_EP_POS = 15;

// one Gibbs call in flight on the device ("job"); pointers are device pointers
struct JobDev {
    int32_t R;           // reads
    int32_t first_read;  // first_read_for_gibbs_initialization
    int32_t n_dense;     // reads stored as dense columns
    int32_t pad0;
    // big state, owned by the slot the job runs in
    double* alpha;  // [NH][T][Kp]
    double* beta;   // [NH][T][Kp]
    double* eG;     // [NH][T][Kp]   eMatGrid_t
    double* c;      // [NH][T]
    uint32_t* W;    // [T][Kp]       32-SNP allele words of the selected haplotypes on this call's SNP axis
    uint32_t* Wc;   // [Tc][Kp]      scratch: words on the panel's common-SNP axis (all-SNP calls only)
    ReadDesc* desc;  // [R]
    TabEnt* tabs;    // table pool
    TabEnt* dense;   // [n_dense][Kp] {E, 1/E} columns of the reads kept dense
    double* xprob;   // [R][4] label probabilities of the last visit (H_class is derived from them): normalised (slot 3 = -1)
                     // or the raw products of the fast decision path (slot 3 = label the read had)
    uint8_t* snp_type;  // [nSNPs] all-SNP calls: 0 common, 1 rare without carrier, 2 rare with carrier(s)
    double* rate;       // [T] scratch (block definition)
    // inputs
    const int32_t* which;  // [K] 1-based
    const int32_t* rs;     // [T + 1] first read of each grid (reads are sorted by wif0)
    const int32_t* roff;   // [R + 1]
    const int32_t* u;      // [nU] 0-based SNP index on this call's axis
    const double* pRA;     // [nU][2] (pR, pA) running values per read-SNP, computed on the host with libm pow()
    const int32_t* wif0;   // [R]
    const int32_t* ts;     // [T + 1] first table-pool entry of each grid's reads (table-mode reads only)
    const int32_t* ginfo;  // [T + 1][4] per grid {rs, ts, reads in its first staging chunk, table end of that chunk}
    const int32_t* dense_reads;  // [n_dense] read indices stored as dense columns
    // haplotype classes of the sweep kernel (k_build_classes; diploid, one CTA per job): per grid the K selected haplotypes grouped by
    // the allele bits the grid's reads can see (own 32-SNP word + the neighbour bits of reads that start / end outside it)
    int32_t* cinfo;      // [T + 1] classes of the grid, 0 = the grid's reads walk all K states
    uint16_t* cperm;     // [T][KA] haplotypes sorted by (class, k); bit 15 = first of its class
    uint8_t* ccls;       // [T][NT][EPT] class of haplotype tid + i * NT at [tid][i]
    uint8_t* cent;       // [T][NT] 1 + class of sorted position EPT * tid - 1 (0 for tid 0)
    uint4* crec;         // [T][CLS_LANES * 32] per class {word g-1 & mask, word g, word g+1 & mask, first sorted position | members << 16}
    const double* runif_reads;  // [n_its][R]
    const double* runif_shard;  // [n_ep][T - 1]
    const double* tm;           // [T - 1][2] (sigma, 1 - sigma)
    const int32_t* H0;          // [R] starting labels
    // three-haplotype (NIPT) block Gibbs only
    const double* runif_block;    // [n_ep][R]
    const double* runif_H_class;  // [n_ep][R]
    const int32_t* L_grid;        // [T]
    int32_t ep_stream;            // 1: runif_block holds the caller's flat episode stream, walked with the position in lik[LIK_EP_POS]
    int32_t pad1;
    unsigned char* blk;           // per-slot scratch of the block definition / resampler (layout: BlockScratch)
    // outputs / evolving small state
    int32_t* H;       // [R] 1-based labels
    int32_t* Hclass;  // [R]
    double* lik;      // [max(n_its,1)][LIK_N]
    int32_t* underflow;  // [1]
    double* hapProbs;    // [3][nSNPs] running sums over sampling sweeps, scaled at the end
    double* genM;        // [3][nSNPs]
    double* genF;        // [3][nSNPs]
    double* hapLocal;    // [3][nSNPs] rare/common calls: the reference's never re-zeroed hapProbs_t_local (gibbs-small.cpp:711-867)
    int32_t* cat_out;    // [R] read_category export
    int32_t* Hs;         // [n_sample][R] labels after each sampling sweep (n_sample > 1 only)
};

struct PanelDev {
    int32_t K_full, Tc, nSNPsC, nMaxDH, n_special, nSNPs_all;
    const uint8_t* hapMatcherR;    // [Tc][K_full]
    const int32_t* distinctHapsB;  // [Tc][nMaxDH]
    const int32_t* special;        // [2][n_special] column-major
    const int32_t* helper;         // [2][Tc] column-major
    const uint8_t* snp_is_common;
    const int32_t* common_snp_index;
    const int64_t* rare_off;
    const int32_t* rare_snps;
    const int8_t* asm_src;         // [T_all][32] all-SNP grid G, bit b: -1 rare, else (common word select << 5) | source bit
    const int32_t* asm_cg0;        // [T_all] first common-axis grid an all-SNP grid draws from
    const uint16_t* hap_perm;      // [Tc][K_full] haplotypes sorted by their symbol at the grid (stable) — full-panel pass only
    const int32_t* hap_symoff;     // [Tc][nMaxDH + 2] segment offsets of hap_perm per symbol
    const int32_t* n_used;         // [Tc] rows of distinctHapsB in use per grid (= max symbol of hapMatcherR[, g])
    double ref_error;
};

// scratch of the NIPT block Gibbs episode, carved out of JobDev::blk (all arrays have T entries unless noted)
struct BlockScratch {
    double* rate2;      // sigma-weighted switch rate per grid boundary
    double* smoothed;   // smoothed rate
    double* lcs;        // [T][9] log c of the nine forward vectors (slot h, emission label i)
    double* logc;       // [3][T] log c_h[g] workspace
    int32_t* idx_a;     // merge-sort index buffers
    int32_t* idx_b;
    int32_t* to_keep;
    int32_t* blocked_grid;
    int32_t* grid_start;
    int32_t* grid_end;
    int32_t* reads_start;
    int32_t* reads_end;
    int32_t* grid_where;
    int32_t* rmflag;    // blocks without reads
    uint8_t* available;
    int32_t* n_blocks;  // [1]
    __host__ __device__ static size_t bytes(int T) { return (size_t)T * (8 * (2 + 9 + 3) + 4 * 10 + 1) + 64 + 16 * 16; }
    __host__ __device__ void carve(unsigned char* p, int T) {
        auto take = [&](size_t n) {
            unsigned char* q = p;
            p += (n + 15) & ~(size_t)15;
            return q;
        };
        rate2 = (double*)take((size_t)T * 8);
        smoothed = (double*)take((size_t)T * 8);
        lcs = (double*)take((size_t)T * 72);
        logc = (double*)take((size_t)T * 24);
        idx_a = (int32_t*)take((size_t)T * 4);
        idx_b = (int32_t*)take((size_t)T * 4);
        to_keep = (int32_t*)take((size_t)T * 4);
        blocked_grid = (int32_t*)take((size_t)T * 4);
        grid_start = (int32_t*)take((size_t)T * 4);
        grid_end = (int32_t*)take((size_t)T * 4);
        reads_start = (int32_t*)take((size_t)T * 4);
        reads_end = (int32_t*)take((size_t)T * 4);
        grid_where = (int32_t*)take((size_t)T * 4);
        rmflag = (int32_t*)take((size_t)T * 4);
        available = (uint8_t*)take((size_t)T);
        n_blocks = (int32_t*)take(64);
    }
};

struct BatchParams {
    int32_t K, Kp, T, NH, nSNPs, n_its, n_burn;
    uint32_t flags;
    double ff;
    double one_over_K;
    double d2;  // 1 / maxDifferenceBetweenReads
    double class_sum_cutoff;
    double prior[3];
    double rlc[7][3];
    double ref_error;
    int32_t rare_common;
    int32_t Jmax;
    uint32_t dbg;  // experiment switches (env QUILT_B200_DBG; 0 in production)
    int32_t cls_min_reads;  // the sweep decides a grid's reads on haplotype-class totals when at least this many are visited
    // NIPT block Gibbs (gibbs-nipt-block.cpp): host-evaluated libm constants so that the scores use the reference's values
    int32_t shuffle_bin_radius;
    double block_q;          // block_gibbs_quantile_prob
    double lhc[6];           // log terms of rcpp_get_log_p_H_class2 for n1..n6 (gibbs-nipt-block.cpp:169-208)
};

}  // namespace qb
