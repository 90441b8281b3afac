/*
 * quilt_b200.h — C ABI of the B200-native QUILT2 per-sample imputation hot path.
 *
 * Drop-in boundary: everything here is plain C (pointers + sizes, no R / Rcpp /
 * torch types).  A thin Rcpp translation unit (shim/quilt_gpu_shim.cpp, built
 * only where R exists) keeps the reference's 63-SEXP entry point and forwards
 * to these functions; INTEGRATION.md shows the binding.
 *
 * Reference interfaces replaced (paths relative to the QUILT repository):
 *   quilt_gpu_gibbs / _batch      <- rcpp_forwardBackwardGibbsNIPT
 *                                    QUILT/src/gibbs-nipt.cpp:2395-3307,
 *                                    .Call glue QUILT/src/RcppExports.cpp:966-1038,
 *                                    R wrapper QUILT/R/RcppExports.R:215-217,
 *                                    production caller QUILT/R/functions.R:2614-2678
 *   quilt_gpu_make_eMatRead_t     <- Rcpp_make_eMatRead_t_for_gibbs_using_objects
 *                                    QUILT/src/gibbs-small.cpp:116-265 and the
 *                                    rare/common sibling :275-465
 *   quilt_gpu_unpack_panel        <- rcpp_int_expand / inflate_fhb /
 *                                    rcpp_simple_binary_matrix_search
 *                                    QUILT/src/copied-from-stitch.cpp:50-108,
 *                                    QUILT/src/gibbs-small.cpp:69-105
 *   quilt_gpu_forward_backward    <- Rcpp_run_forward_haploid / Rcpp_run_backward_haploid
 *                                    QUILT/src/copied-from-stitch.cpp:340-409
 *   quilt_gpu_haploid_dosage_versus_refs / _batch
 *                                 <- Rcpp_haploid_dosage_versus_refs (full-panel haploid Li-Stephens pass)
 *                                    QUILT/src/reference-single.cpp:2189-2413 (forward :878-1131, backward :1781-2179,
 *                                    eMatDH :272-329, top matches :196-266), .Call seam QUILT/R/functions.R:2034-2070
 *   quilt_gpu_select_haps         <- select_new_haps_mspbwt_v3 (heuristic_approach "A")
 *                                    QUILT/R/mspbwt.R:230-474, called at QUILT/R/functions.R:856-868
 *
 * All real arrays are fp64, column-major (R / Armadillo), K is the fastest
 * dimension.  Integer arrays are int32 unless stated.  Indices are 0-based
 * unless the field name says otherwise (the reference mixes both; the names
 * below keep the reference's convention so the shim is a flat copy).
 *
 * The functions never generate randomness: every uniform the reference draws
 * from R's RNG inside the .Call is an input ("RNG").
 *
 * Supported argument space of the GPU library (everything else returns
 * QUILT_ERR_UNSUPPORTED / QUILT_ERR_BAD_ARG, never a CPU computation):
 *   diploid calls (QUILT_F_SAMPLE_IS_DIPLOID, ff == 0)  K <= 8192
 *   NIPT calls (three haplotypes, 0 <= ff < 1)           K <= 2048, no shard pass
 *   32 SNPs per grid; shard pass with QUILT_F_SHARD_CHECK_EVERY_PAIR.
 *
 * Dependency of the diploid path on the reference's build: for two haplotypes the block resampler
 * (gibbs-nipt-block.cpp:1636-1967) is skipped because every permutation score of the reference is NaN there — the
 * third columns of its local matrices and c3 are ZERO-FILLED by Armadillo's constructors (Armadillo >= 10.5; the
 * compiled reference in oracle/_ref confirms the state is left unchanged).  Against an Armadillo older than 10.5
 * the reference reads uninitialised memory at that point and no implementation can match it.
 */
#ifndef QUILT_B200_H
#define QUILT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- flags (QuiltGibbsArgs.flags); names follow param_list, functions.R:2566-2599 */
#define QUILT_F_SAMPLE_IS_DIPLOID            (1u << 0)
#define QUILT_F_GIBBS_INITIALIZE_ITERATIVELY (1u << 1)
#define QUILT_F_PERFORM_BLOCK_GIBBS          (1u << 2)
#define QUILT_F_DO_SHARD_BLOCK_GIBBS         (1u << 3)
#define QUILT_F_SHARD_CHECK_EVERY_PAIR       (1u << 4)
#define QUILT_F_DISABLE_READ_CATEGORY_USAGE  (1u << 5)
#define QUILT_F_FORCE_RESET_READ_CATEGORY_0  (1u << 6)
#define QUILT_F_MAKE_EMATREAD_RARE_COMMON    (1u << 7)
#define QUILT_F_RESCALE_EMATREAD             (1u << 8)
#define QUILT_F_RECORD_READ_SET              (1u << 9)
#define QUILT_F_USE_SMOOTH_CM_IN_BLOCK_GIBBS (1u << 10)
#define QUILT_F_RETURN_ALPHA                 (1u << 11)  /* debug: copy alpha/beta/c/eMatGrid out */
#define QUILT_F_RETURN_EXTRA                 (1u << 12)  /* debug: copy eMatRead_t out            */
#define QUILT_F_OUTPUT_NO_PROBS              (1u << 14)  /* hapProbs_t / genProbs*_t stay on the device (intermediate call of a
                                                            device-resident chain, quilt_gpu_batch_chain_select): not copied out */
#define QUILT_F_GIBBS_INITIALIZE_AT_FIRST_READ (1u << 13) /* first_read_for_gibbs_initialization is 0 and is not drawn (gibbs-nipt.cpp:2846) */

/* production flag word for QUILT2 diploid common-SNP calls (functions.R:620-706) */
#define QUILT_FLAGS_QUILT2_DIPLOID \
    (QUILT_F_SAMPLE_IS_DIPLOID | QUILT_F_PERFORM_BLOCK_GIBBS | QUILT_F_DO_SHARD_BLOCK_GIBBS | \
     QUILT_F_SHARD_CHECK_EVERY_PAIR | QUILT_F_RESCALE_EMATREAD | QUILT_F_RECORD_READ_SET |     \
     QUILT_F_USE_SMOOTH_CM_IN_BLOCK_GIBBS)

/* status codes */
#define QUILT_OK                0
#define QUILT_ERR_CUDA          1
#define QUILT_ERR_BAD_ARG       2
#define QUILT_ERR_UNSUPPORTED   3
#define QUILT_ERR_NO_DEVICE     4

/*
 * Prepared reference panel in the compressed form QUILT2 keeps after
 * QUILT_prepare_reference (format pinned by
 * QUILT/tests/testthat/test-unit-reference-single.R:210-309).
 */
typedef struct QuiltPanel {
    int32_t K_full;          /* haplotypes in the panel                               */
    int32_t nGrids;          /* T over the panel's (common) SNPs = ceil(nSNPs/32)     */
    int32_t nSNPs;           /* (common) SNPs                                          */
    int32_t nMaxDH;          /* rows of distinctHapsB / distinctHapsIE (<= 255)        */
    const uint8_t* hapMatcherR;      /* [K_full x nGrids] 1-based row of distinctHapsB, 0 = special */
    const int32_t* distinctHapsB;    /* [nMaxDH x nGrids] packed 32-SNP words, LSB = first SNP      */
    const double*  distinctHapsIE;   /* [nMaxDH x nSNPs]  (1-eps) if bit set else eps                */
    const int32_t* eMatDH_special_matrix;        /* [n_special x 2] col 0 = 0-based hap, col 1 = word */
    int32_t        n_special;
    const int32_t* eMatDH_special_matrix_helper; /* [nGrids x 2] 1-based first / last row per grid    */
    double  ref_error;
    /* rare/common extras for the all-SNP stage (nSNPs_all == 0 when absent), rare_common.R:202-322 */
    int32_t nSNPs_all;
    const uint8_t* snp_is_common;      /* [nSNPs_all]                                      */
    const int32_t* common_snp_index;   /* [nSNPs_all] 1-based index among common SNPs      */
    const int64_t* rare_hap_offsets;   /* [K_full + 1] CSR over rare_per_hap_info          */
    const int32_t* rare_hap_snps;      /* 1-based all-SNP indices where the hap carries alt */
} QuiltPanel;

/* sampleReads flattened to CSR (reference: R list of list(J, wif, bq, u), test-drivers.R:222-227) */
typedef struct QuiltReads {
    int32_t nReads;
    const int32_t* offsets;  /* [nReads + 1]; read r owns u/bq[offsets[r] .. offsets[r+1]) ; J_r = count - 1 */
    const int32_t* u;        /* 0-based SNP index                                        */
    const int32_t* bq;       /* signed phred: <0 ref allele seen, >0 alt allele seen     */
    const int32_t* wif0;     /* [nReads] 0-based grid of the read's central SNP, non-decreasing */
} QuiltReads;

/* one call of rcpp_forwardBackwardGibbsNIPT (one sample, one Gibbs start, S = 1) */
typedef struct QuiltGibbsArgs {
    const QuiltPanel* panel;
    QuiltReads reads;
    int32_t K;                          /* Ksubset                                           */
    const int32_t* which_haps_to_use;   /* [K] 1-based, order defines the k axis             */
    int32_t nGrids;                     /* T of THIS call (common-SNP or all-SNP grids)      */
    int32_t nSNPs;                      /* SNPs of THIS call (= length(grid))                */
    const double*  transMatRate_tc_H;   /* [2 x (nGrids-1)] row 0 = sigma (stay), row 1 = 1-sigma */
    const int32_t* L_grid;              /* [nGrids] physical position per grid               */
    const double*  smooth_cm;           /* [nGrids-1] (unused by results, kept for parity)   */
    double  ff;                         /* fetal fraction; 0 for diploid                     */
    int32_t n_gibbs_burn_in_its;
    int32_t n_gibbs_sample_its;
    const int32_t* block_gibbs_iterations; /* 0-based sweep indices, e.g. {3,6,9}            */
    int32_t n_block_gibbs_iterations;
    const int32_t* H0;                  /* [nReads] starting labels, 1-based                 */
    int32_t first_read_for_gibbs_initialization; /* 0-based (gibbs-nipt.cpp:2848)            */
    const double* runif_reads;          /* [nReads * n_full_its]      (gibbs-nipt.cpp:2845)  */
    const double* runif_block;          /* [n_block_its x nReads]     (gibbs-nipt.cpp:3017)  */
    const double* runif_shard;          /* [n_block_its x (nGrids-1)] (gibbs-nipt-block.cpp:2054) */
    const double* runif_H_class;        /* NIPT only: [n_block_its x nReads] unif_rand per Rcpp::sample (block.cpp:226-243) */
    double  maxDifferenceBetweenReads;
    int32_t Jmax;
    double  class_sum_cutoff;
    int32_t shuffle_bin_radius;
    double  block_gibbs_quantile_prob;
    uint32_t flags;
    /* Episode stream (optional, replaces runif_block / runif_shard / runif_H_class when non-NULL): R's unif_rand()
     * stream as it stands right after the call's first two draws (runif(nReads * n_full_its), sample(nReads, 1);
     * gibbs-nipt.cpp:2845-2848).  The library consumes it in the reference's own order — per block-Gibbs episode
     * 6 x nReads (runif_proposed, drawn but unused), nReads (runif_block), nReads (runif_total, unused)
     * (gibbs-nipt.cpp:3013-3018), then for NIPT one value per read whose H_class is 0/4/5/6/7
     * (rcpp_sample_H_using_H_class, gibbs-nipt-block.cpp:213-246: a DATA-DEPENDENT count), then for the diploid
     * shard pass nGrids - 1 values (gibbs-nipt-block.cpp:2054) — and reports in QuiltGibbsOut.n_unif_consumed how
     * many values the reference would have drawn (episodes after an underflow early return draw nothing,
     * gibbs-nipt.cpp:2959-2969), so that the caller can leave R's generator at exactly the reference's position.
     * n_unif_stream >= n_episodes * (8 * nReads + (NIPT ? nReads : 0) + (shard ? nGrids - 1 : 0)).              */
    const double* unif_stream;
    int64_t n_unif_stream;
} QuiltGibbsArgs;

typedef struct QuiltGibbsOut {
    int32_t underflow_problem;   /* reference: list(underflow_problem=TRUE) early return     */
    double* hapProbs_t;          /* [3 x nSNPs]                                              */
    double* genProbsM_t;         /* [3 x nSNPs]                                              */
    double* genProbsF_t;         /* [3 x nSNPs]                                              */
    int32_t* H;                  /* [nReads] final labels                                    */
    int32_t* H_class;            /* [nReads]                                                 */
    double* per_it_likelihoods;  /* [n_full_its x 13] column-major                           */
    /* optional (may be NULL); filled when QUILT_F_RETURN_ALPHA / _EXTRA are set */
    double* alphaHat_t[3];       /* each [K x nGrids]                                        */
    double* betaHat_t[3];
    double* eMatGrid_t[3];
    double* c[3];                /* each [nGrids]                                            */
    double* eMatRead_t;          /* [K x nReads]                                             */
    int32_t* read_category;      /* [nReads]                                                 */
    int32_t* H_sample_its;       /* optional [nReads x n_gibbs_sample_its]: the labels after each sampling sweep — the reference's
                                    double_list_of_ending_read_labels[[1]][[i]] (gibbs-nipt.cpp:3104); column n_sample_its - 1 == H */
    /* filled by every call */
    int64_t n_unif_consumed;     /* values of unif_stream the reference would have drawn (0 without unif_stream) */
    int32_t underflow_iteration; /* 0-based sweep whose underflow check failed (gibbs-nipt.cpp:2959-2969), -1 if none */
} QuiltGibbsOut;

/* ------------------------------------------------------------------ GPU library (libquiltgpu.so) */

/* drop-in granularity: one sample, one call; host pointers in, host pointers out */
int quilt_gpu_gibbs(const QuiltGibbsArgs* args, QuiltGibbsOut* out);
/* benchmark / production granularity: n independent samples, one launch sequence */
int quilt_gpu_gibbs_batch(int32_t n, const QuiltGibbsArgs* args, QuiltGibbsOut* out);

/* staged form of the same call, so inputs can be resident in HBM before a timed region */
typedef struct QuiltGpuBatch QuiltGpuBatch;
int quilt_gpu_batch_stage(int32_t n, const QuiltGibbsArgs* args, QuiltGpuBatch** batch); /* H2D   */
int quilt_gpu_batch_run(QuiltGpuBatch* batch);                                           /* kernels, async on the library stream */
int quilt_gpu_batch_sync(QuiltGpuBatch* batch);
int quilt_gpu_batch_fetch(QuiltGpuBatch* batch, QuiltGibbsOut* out);                     /* D2H   */
int quilt_gpu_batch_free(QuiltGpuBatch* batch);
/* device time of the last quilt_gpu_batch_run in ms, measured with CUDA events on the library stream;
 * sweep_ms = time inside the dominant (sweep) kernel launches, n_sweep_launches their count */
int quilt_gpu_batch_timing(QuiltGpuBatch* batch, double* total_ms, double* sweep_ms, int32_t* n_sweep_launches);
/* bytes moved by quilt_gpu_batch_stage (host -> device) and quilt_gpu_batch_fetch (device -> host), and the ALGORITHMIC bytes of
 * the sweep kernel launches of one run: sum over launches and jobs of 8 * K * (5 * nHap * nGrids + nReads) */
int quilt_gpu_batch_bytes(QuiltGpuBatch* batch, int64_t* h2d_bytes, int64_t* d2h_bytes, double* sweep_algorithmic_bytes);

/*
 * Haplotype re-selection between Gibbs calls (select_new_haps_mspbwt_v3, QUILT/R/mspbwt.R:230-474).
 *
 * In-tree part, restated exactly: hap = round(hapProbs_t[h, ]) packed 32 SNPs per word (rcpp_int_contract, :277-278);
 * the nIndices interleaved grid subsets seq(i, nGrids, nIndices) (:283); per subset the match table re-keyed to
 * (index1, start1, end1, len1), ordered by (index1, -end1, -start1) and stripped of adjacent rows repeating
 * (index1, start1) (:312-335); subsets concatenated and ordered by -len1 (:340-349); then, if the haplotypes found
 * exceed Knew, the coverage-weighted ranking (:414-441) and the interleave / unique / first-Knew cut (:443-466);
 * otherwise the haplotypes found in order of first appearance (:368-379).  Padding a short list with sample()
 * (:381-399) draws from R's generator and stays with the caller: n_found < Knew tells it how many to add.
 *
 * Un-vendored part (mspbwt 0.1.0: map_Z_to_all_symbols, Rcpp_find_good_matches_without_a; parity UNPINNED — the
 * package is not in the reference tree): its contract is re-derived from the call site
 * (:284-310).  A word of Z maps to the row of distinctHapsB[, g] holding the same 32-SNP word (hapMatcherR's symbol);
 * a word that is not in the panel's table matches no haplotype, and symbol 0 ("special") never matches.  Walking a
 * subset left to right, every panel haplotype carries the length of its current run of identical symbols with Z;
 * at a position the 2 * mspbwtL haplotypes with the longest runs (ties: lower haplotype index) are Z's neighbours;
 * a run is reported as (start0, index0, len1) when it ENDS — next symbol differs or the subset ends — while its
 * haplotype is a neighbour and len1 >= mspbwtM.
 */
typedef struct QuiltSelectArgs {
    const QuiltPanel* panel;     /* hapMatcherR / distinctHapsB of the common-SNP panel                     */
    int32_t nHap;                /* haplotypes of the sample: 2 (diploid) or 3 (nipt)                       */
    const double* hapProbs_t;    /* [3 x nSNPs] column-major, as returned by the Gibbs call                 */
    int32_t Knew;
    int32_t mspbwt_nindices;     /* 4 (quilt.R:174)                                                         */
    int32_t mspbwtL;             /* 3 (quilt.R:171)                                                         */
    int32_t mspbwtM;             /* 1 (quilt.R:172)                                                         */
} QuiltSelectArgs;
/* which_haps_to_use [Knew] 1-based, first n_found entries valid; n_unique = haplotypes found before the cut */
int quilt_gpu_select_haps(const QuiltSelectArgs* args, int32_t* which_haps_to_use, int32_t* n_found, int32_t* n_unique);
/* Device-resident chaining: job j of `next` (staged, not yet run) follows job j of `prev` (run): select from prev's
 * hapProbs_t on the device and write the list into next's which_haps_to_use, no host round trip.  A list shorter than
 * next's Ksubset is completed from haplotypes not yet in it by a partial Fisher-Yates driven by pad_unif
 * ([n_jobs x Ksubset] uniforms; statistically what sample() does at mspbwt.R:381-399, not R's bit stream).        */
int quilt_gpu_batch_chain_select(QuiltGpuBatch* prev, QuiltGpuBatch* next, int32_t mspbwt_nindices, int32_t mspbwtL, int32_t mspbwtM,
                                 const double* pad_unif);
/* The same link with HOST buffers: quilt_gpu_gibbs_batch (waves pipelined: H2D / kernels / D2H overlap) whose calls take their
 * which_haps_to_use from the device-resident results of `prev` (NULL: the lists in args are used).  With `kept` non-NULL the
 * batch object survives the call (results resident on the device) so that it can be `prev` of the next link; free it with
 * quilt_gpu_batch_free. */
int quilt_gpu_gibbs_batch_chained(int32_t n, const QuiltGibbsArgs* args, QuiltGibbsOut* out, QuiltGpuBatch* prev, int32_t mspbwt_nindices,
                                  int32_t mspbwtL, int32_t mspbwtM, const double* pad_unif, QuiltGpuBatch** kept);
/* The whole chain in ONE call: stage s holds call s of every (sample, chain) pair (args[s][0 .. n), out[s][0 .. n)); from stage 1 on
 * the lists come from stage s - 1 on the device (pad_unif[s - 1]: [n x Ksubset]).  A single wave pipeline spans the stages: the
 * host prepares stage s + 1 and unpacks stage s while the GPU computes. */
int quilt_gpu_gibbs_chain(int32_t n_stages, int32_t n, const QuiltGibbsArgs* const* args, QuiltGibbsOut* const* out, int32_t mspbwt_nindices,
                          int32_t mspbwtL, int32_t mspbwtM, const double* const* pad_unif);
/* device time (ms, CUDA events on the library stream) of the chained selection that filled this batch's lists */
int quilt_gpu_batch_chain_timing(QuiltGpuBatch* batch, double* chain_ms);
/* the haplotype list job `job` of a staged batch currently holds on the device (after a chained selection: the new list) */
int quilt_gpu_batch_which_haps(QuiltGpuBatch* batch, int32_t job, int32_t* which_haps_to_use /*[Ksubset]*/);

/*
 * Full-panel haploid pass (QUILT1 / use_mspbwt = FALSE, truth-haplotype diagnostics): one haplotype's genotype likelihoods
 * against ALL K_full panel haplotypes, per-grid emissions looked up from the (nMaxDH + 1)-entry table eMatDH by the
 * hapMatcherR byte, "special" haplotypes (symbol 0) recomputed from their bits, lazy normalisation of the forward
 * variables (only when the running minimum emission drops below the threshold), dosage through per-symbol gamma sums,
 * best-matching haplotypes at the thinned grids.  Production settings (functions.R:2034-2070): use_eMatDH, hapMatcherR,
 * special symbols, always_normalize = FALSE, normalize_emissions = TRUE, the "version 3" kernels.
 */
#define QUILT_HF_RETURN_DOSAGE   (1u << 0)
#define QUILT_HF_RETURN_BETAHAT  (1u << 1)
#define QUILT_HF_RETURN_GAMMA    (1u << 2)
#define QUILT_HF_GET_BEST_HAPS   (1u << 3)   /* get_best_haps_from_thinned_sites */
#define QUILT_HF_RETURN_ALPHAHAT (1u << 4)   /* (R passes full_alphaHat_t as scratch and may read it back) */
typedef struct QuiltHaploidArgs {
    const QuiltPanel* panel;
    const double* gl;                        /* [2 x nSNPs] (ref, alt) likelihoods per SNP: make_gl_from_u_bq, reference-single.R:19-42 */
    const double* transMatRate_t;            /* [2 x (nGrids - 1)] row 0 = no recombination, row 1 = recombination              */
    const int32_t* gammaSmall_cols_to_get;   /* [nGrids] -1, or the 0-based thinned column collecting best haplotypes (quilt.R:719-721) */
    int32_t n_thinned;                       /* number of thinned columns (max of the above + 1)                                 */
    int32_t K_top_matches;                   /* 5 (quilt.R:115)                                                                  */
    int32_t best_cap;                        /* capacity per thinned column of the best-haplotype outputs (ties can exceed K_top_matches) */
    double  min_emission_prob_normalization_threshold;   /* 1e-100                                                                */
    uint32_t flags;
} QuiltHaploidArgs;
typedef struct QuiltHaploidOut {
    double* dosage;              /* [nSNPs]                 (QUILT_HF_RETURN_DOSAGE)                                   */
    double* c;                   /* [nGrids]                                                                            */
    double* alphaHat_t;          /* [K_full x nGrids]       (QUILT_HF_RETURN_ALPHAHAT) lazily normalised forward variables */
    double* betaHat_t;           /* [K_full x nGrids]       (QUILT_HF_RETURN_BETAHAT)                                  */
    double* gamma_t;             /* [K_full x nGrids]       (QUILT_HF_RETURN_GAMMA)                                    */
    int32_t* best_haps;          /* [n_thinned x best_cap] 0-based haplotypes with gamma >= the K_top_matches-th largest, in haplotype order */
    double*  best_haps_values;   /* [n_thinned x best_cap] their gamma                                                 */
    int32_t* best_haps_count;    /* [n_thinned] how many qualified (entries beyond best_cap are dropped)               */
} QuiltHaploidOut;
int quilt_gpu_haploid_dosage_versus_refs(const QuiltHaploidArgs* args, QuiltHaploidOut* out);
/* n independent passes over one panel (the haplotypes of many samples / chains): one CTA per pass */
int quilt_gpu_haploid_dosage_versus_refs_batch(int32_t n, const QuiltHaploidArgs* args, QuiltHaploidOut* out);
/* device time (ms) of the last batch call's kernels and the ALGORITHMIC bytes they stand for: per pass
 * K_full * nGrids * (2 x 1 symbol byte + 8 alphaHat written + 8 alphaHat read) (+ 8 per returned matrix element) */
int quilt_gpu_haploid_last_timing(double* kernel_ms, double* algorithmic_bytes);

/*
 * Per-sample reductions that follow the Gibbs calls (SURVEY.md section 8 row a12 and the INFO counters of section 8e), on the
 * device from the device-resident hapProbs_t of staged / kept batches, so that only per-sample vectors travel:
 *   dosage / gp_t running sums over the stored calls in R's order and the final division (QUILT/R/functions.R:999-1020,
 *   :1100-1122, :1304-1325); recast_haps of the phasing call against the accumulated gp_t (:3180-3209, called at :1211 /
 *   :1236) -> recast haplotype dosages and the phased GT; eij / fij / max_gen (:1408-1411) and the per-rank counters
 *   infoCount, afCount, hweCount (QUILT/R/quilt.R:957-961) summed over the samples in order.  Diploid samples.
 *   (R's round(x, 3) is taken as rint(1000 x) / 1000.)
 */
typedef struct QuiltSummaryCall {
    QuiltGpuBatch* batch;        /* a batch that has been run (staged form or kept by quilt_gpu_gibbs_batch_chained) */
    int32_t job;                 /* index of the call inside it                                                    */
} QuiltSummaryCall;
typedef struct QuiltSampleSummary {
    int32_t n_calls;             /* stored calls of the sample, in the reference's accumulation order            */
    const QuiltSummaryCall* calls;
    QuiltSummaryCall phasing;    /* the phasing chain's final call (phasing_haps)                                 */
    double* dosage;              /* [nSNPs]      out                                                              */
    double* gp_t;                /* [3 x nSNPs]  out                                                              */
    double* hd;                  /* [nSNPs x 2]  out: recast haplotype dosages, column-major                      */
    int8_t* gt;                  /* [nSNPs x 2]  out: round(hd), the phased genotype                              */
} QuiltSampleSummary;
int quilt_gpu_samples_summary(int32_t n_samples, int32_t nSNPs, const QuiltSampleSummary* samples,
                              double* infoCount /*[nSNPs x 2] or NULL*/, double* afCount /*[nSNPs]*/, double* hweCount /*[nSNPs x 3]*/);

/*
 * The steps on either side of the path.
 *
 * quilt_gpu_ingest_pileup: one sample's flat pileup (what loadBamAndConvert leaves in sampleReads, QUILT/R/functions.R:243-272:
 * per read the 0-based SNP indices u, the signed scaled base qualities bq and the 0-based central SNP) -> the reads in the
 * order the Gibbs path needs them: wif0 = grid[central SNP] (snap_sampleReads_to_grid, :295-298), reads ordered by wif0 (stable),
 * the first read of every grid, grid_has_read (:314-316), and the sample's allele counts (get_alleleCount, :2779-2800:
 * alleleCount[, 1] = sum of P(alt), alleleCount[, 2] = sum of P(ref) + P(alt) over the read-SNP entries of the SNP, added in
 * entry order like increment2N, QUILT/src/copied-from-stitch.cpp:573-579).  BAM decoding stays on the CPU.
 */
typedef struct QuiltPileup {
    int32_t nReads, nSNPs, nGrids;
    const int32_t* offsets;      /* [nReads + 1]                                    */
    const int32_t* u;            /* [offsets[nReads]] 0-based SNP index             */
    const int32_t* bq;           /* [offsets[nReads]] signed scaled base quality    */
    const int32_t* central_snp;  /* [nReads] 0-based central SNP of the read        */
    const int32_t* grid;         /* [nSNPs] 0-based grid of every SNP               */
} QuiltPileup;
typedef struct QuiltIngestOut {   /* any pointer may be NULL */
    int32_t* order;              /* [nReads] pileup index of the read at each position of the path order */
    int32_t* offsets;            /* [nReads + 1] */
    int32_t* u;                  /* [nU] */
    int32_t* bq;                 /* [nU] */
    int32_t* wif0;               /* [nReads] non-decreasing */
    int32_t* first_read_of_grid; /* [nGrids + 1] */
    uint8_t* grid_has_read;      /* [nGrids] */
    double* alleleCount;         /* [nSNPs x 2] column-major */
} QuiltIngestOut;
int quilt_gpu_ingest_pileup(const QuiltPileup* in, QuiltIngestOut* out);

/*
 * quilt_gpu_make_vcf_column: the per-sample VCF column of a diploid sample (QUILT/R/functions.R:1408-1463:
 * STITCH::rcpp_make_column_of_vcf with the GT replaced by the phased genotype), FORMAT GT:GP:DS:HD, one fixed-width record of
 * QUILT_VCF_RECORD bytes per SNP, "a|b:%.3f,%.3f,%.3f:%.3f:%.3f,%.3f" (no terminator).
 */
#define QUILT_VCF_RECORD 39
int quilt_gpu_make_vcf_column(int32_t nSNPs, const double* gp_t /*[3 x nSNPs]*/, const double* hd /*[nSNPs x 2]*/, char* out /*[nSNPs][QUILT_VCF_RECORD]*/);

/* component entry points (parity tests of the individual reference functions) */
int quilt_gpu_make_eMatRead_t(const QuiltGibbsArgs* args, double* eMatRead_t /*[K x nReads]*/,
                              int32_t* read_category /*[nReads] or NULL*/);
int quilt_gpu_unpack_panel(const QuiltPanel* panel, int32_t K, const int32_t* which_haps_to_use,
                           int32_t all_snps /*0: common grids, 1: all-SNP grids*/,
                           uint32_t* words /*[K x nGrids(_all)]*/);
int quilt_gpu_forward_backward(int32_t K, int32_t nGrids, const double* eMatGrid_t, const double* transMatRate_tc_H,
                               double* alphaHat_t, double* betaHat_t, double* c);

/* housekeeping */
int         quilt_gpu_device_count(void);
int         quilt_gpu_set_device(int32_t device);
const char* quilt_gpu_last_error(void);
int64_t     quilt_gpu_kernel_launches(void);   /* cumulative count of this library's kernel launches */
/* Per-section timing: the reference's print_extra_timing_information / suppressOutput = 0 switch (QUILT/src/copied-from-stitch.cpp:31-45
 * prints the time spent between code sections).  enable != 0: from now on every kernel launch of the library is followed by a CUDA
 * event; quilt_gpu_section_report writes a table "section, launches, total ms, avg ms, share" (device time between consecutive
 * events on the library stream, per kernel) into buf (NUL-terminated, truncated to cap) and returns the size the full text needs.
 * enable == 0 switches it off and drops the events.  Off by default (no events are created). */
int         quilt_gpu_section_timing(int32_t enable);
int64_t     quilt_gpu_section_report(char* buf, int64_t cap);
void        quilt_gpu_release_panel_cache(void);

#ifdef __cplusplus
}
#endif
#endif /* QUILT_B200_H */
