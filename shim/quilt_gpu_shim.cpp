// quilt_gpu_shim.cpp — Rcpp translation unit that keeps the reference's .Call entry point for the Gibbs hot path
// and forwards it to the C ABI of libquiltgpu.so (include/quilt_b200.h).
//
// It replaces, symbol for symbol,
//     RcppExport SEXP _QUILT_rcpp_forwardBackwardGibbsNIPT(SEXP x 63)      QUILT/src/RcppExports.cpp:966-1038
// i.e. the glue of rcpp_forwardBackwardGibbsNIPT (QUILT/src/gibbs-nipt.cpp:2395-3307), so
// QUILT/R/RcppExports.R:215-217 and the production caller QUILT/R/functions.R:2614-2678 stay untouched
// (INTEGRATION.md shows the Makevars change).  Where R + Rcpp exist it is built against them; in this repository's
// image it is compiled, linked and EXECUTED against the header-only stand-in of oracle/refshim/ with the CPU oracle
// as back end (tests/test_shim_executes.py: mock R list in -> C ABI -> named list out, compared field by field with
// the compiled reference, including the position the random stream is left at).
//
// What the shim does, in the reference's order:
//   1. rejects every argument / param_list flag outside the accelerated space with Rcpp::stop (never a silent CPU path);
//   2. flattens sampleReads (R list of list(J, wif, bq, u), gibbs-small.cpp:149-152) into CSR arrays;
//   3. draws from R's generator exactly what the reference draws, in its order (SURVEY.md section 8b "RNG"):
//      runif(nReads * n_full_its) (:2845), sample(nReads, 1) (:2848, unless gibbs_initialize_at_first_read),
//      runif(nReads) for the labels when use_starting_read_labels = FALSE (:2862).  What follows depends on the DATA
//      (per block-Gibbs episode 8 x runif(nReads) :3013-3018; for NIPT one unif_rand() per read whose H_class is
//      0/4/5/6/7, gibbs-nipt-block.cpp:213-246; for diploid samples runif(nGrids - 1), :2054; and nothing at all after
//      an underflow early return, gibbs-nipt.cpp:2959-2969), so the shim marks the generator's state, hands the library
//      the longest stream the call could need (QuiltGibbsArgs.unif_stream) and afterwards rewinds to the mark and
//      advances by the number of values the library reports as consumed (QuiltGibbsOut.n_unif_consumed): R's stream is
//      left exactly where the reference would leave it, also for NIPT and after an underflow retry;
//   4. calls the back end (host pointers in / out, no R types);
//   5. rebuilds the named return list of gibbs-nipt.cpp:3217-3306 for the production param_list.
// The scratch matrices R passes in (alphaHat_t*, betaHat_t*, eMatGrid_t*, gamma*_t_local, eMatRead_t) are ignored:
// the device owns that state (R never reads them back in production, functions.R:2594 / quilt.R:731-762).
#include <Rcpp.h>

#include <cstring>
#include <string>
#include <vector>

#include "../include/quilt_b200.h"

// back end: the CUDA library; the CPU test build points these at the oracle (-DQUILT_SHIM_BACKEND=quilt_oracle_gibbs)
#ifndef QUILT_SHIM_BACKEND
#define QUILT_SHIM_BACKEND quilt_gpu_gibbs
#define QUILT_SHIM_LAST_ERROR() quilt_gpu_last_error()
#else
extern "C" int QUILT_SHIM_BACKEND(const QuiltGibbsArgs*, QuiltGibbsOut*);
#define QUILT_SHIM_LAST_ERROR() "(CPU test back end)"
#endif

using namespace Rcpp;

namespace {

inline bool flag(const List& pl, const char* name) { return as<bool>(pl[name]); }

// ---- R's generator: mark / rewind.  Real R keeps the state in .Random.seed (written by PutRNGstate, read by
// GetRNGstate); the stand-in exposes the same two operations on its injected stream.
#ifdef REFSHIM_RCPPARMADILLO_H
struct RngMark {
    long long pos;
};
inline RngMark rng_mark() { return RngMark{refshim::rng().save()}; }
inline void rng_rewind(const RngMark& m) { refshim::rng().restore(m.pos); }
#else
struct RngMark {
    Rcpp::RObject seed;
};
inline RngMark rng_mark() {
    PutRNGstate();  // flush the generator's state into .Random.seed
    RngMark m;
    m.seed = Rf_duplicate(Rf_findVarInFrame(R_GlobalEnv, Rf_install(".Random.seed")));
    GetRNGstate();
    return m;
}
inline void rng_rewind(const RngMark& m) {
    Rf_defineVar(Rf_install(".Random.seed"), Rf_duplicate(m.seed), R_GlobalEnv);
    GetRNGstate();  // reload the generator from .Random.seed
}
#endif

}  // namespace

// [[Rcpp::export]]
RcppExport SEXP _QUILT_rcpp_forwardBackwardGibbsNIPT(
    SEXP sampleReadsSEXP, SEXP eMatRead_tSEXP, SEXP priorCurrent_mSEXP, SEXP alphaMatCurrent_tcSEXP, SEXP eHapsCurrent_tcSEXP,
    SEXP transMatRate_tc_HSEXP, SEXP ffSEXP, SEXP blocks_for_outputSEXP, SEXP alphaHat_t1SEXP, SEXP betaHat_t1SEXP, SEXP alphaHat_t2SEXP,
    SEXP betaHat_t2SEXP, SEXP alphaHat_t3SEXP, SEXP betaHat_t3SEXP, SEXP eMatGrid_t1SEXP, SEXP eMatGrid_t2SEXP, SEXP eMatGrid_t3SEXP,
    SEXP gammaMT_t_localSEXP, SEXP gammaMU_t_localSEXP, SEXP gammaP_t_localSEXP, SEXP hapSum_tcSEXP, SEXP hapMatcherSEXP, SEXP hapMatcherRSEXP,
    SEXP use_hapMatcherRSEXP, SEXP distinctHapsBSEXP, SEXP distinctHapsIESEXP, SEXP eMatDH_special_matrix_helperSEXP,
    SEXP eMatDH_special_matrixSEXP, SEXP rhb_tSEXP, SEXP ref_errorSEXP, SEXP which_haps_to_useSEXP, SEXP wif0SEXP, SEXP grid_has_readSEXP,
    SEXP L_gridSEXP, SEXP smooth_cmSEXP, SEXP param_listSEXP, SEXP skip_read_iterationSEXP, SEXP Jmax_localSEXP,
    SEXP maxDifferenceBetweenReadsSEXP, SEXP maxEmissionMatrixDifferenceSEXP, SEXP run_fb_grid_offsetSEXP, SEXP gridSEXP,
    SEXP snp_start_1_basedSEXP, SEXP snp_end_1_basedSEXP, SEXP generate_fb_snp_offsetsSEXP, SEXP suppressOutputSEXP, SEXP n_gibbs_startsSEXP,
    SEXP n_gibbs_sample_itsSEXP, SEXP n_gibbs_burn_in_itsSEXP, SEXP double_list_of_starting_read_labelsSEXP, SEXP seed_vectorSEXP,
    SEXP prev_list_of_alphaBetaBlocksSEXP, SEXP i_snp_block_for_alpha_betaSEXP, SEXP do_block_resamplingSEXP, SEXP artificial_relabelSEXP,
    SEXP class_sum_cutoffSEXP, SEXP shuffle_bin_radiusSEXP, SEXP block_gibbs_iterationsSEXP, SEXP block_gibbs_quantile_probSEXP,
    SEXP rare_per_hap_infoSEXP, SEXP common_snp_indexSEXP, SEXP snp_is_commonSEXP, SEXP rare_per_snp_infoSEXP) {
    BEGIN_RCPP
    Rcpp::RNGScope rngScope;  // same bracket as the generated glue (RcppExports.cpp:971)
    const List sampleReads(sampleReadsSEXP);
    const List pl(param_listSEXP);
    const int nReads = sampleReads.size();
    const NumericVector tm(transMatRate_tc_HSEXP);  // cube [2 x (nGrids - 1) x 1]
    const IntegerVector tm_dim = tm.attr("dim");
    if (tm_dim.size() < 2 || tm_dim[0] != 2) stop("quilt-b200: transMatRate_tc_H must be a [2 x (nGrids - 1) x 1] array");
    if (tm_dim.size() >= 3 && tm_dim[2] != 1) stop("quilt-b200: S (number of parameter sets) must be 1");
    const int nGrids = tm_dim[1] + 1;
    const IntegerVector grid(gridSEXP);
    const int nSNPs = grid.size();
    const int n_starts = as<int>(n_gibbs_startsSEXP);
    const int n_sample = as<int>(n_gibbs_sample_itsSEXP), n_burn = as<int>(n_gibbs_burn_in_itsSEXP);
    const int n_full = n_sample + n_burn;
    const double ff = as<double>(ffSEXP);
    const IntegerVector seed_vector(seed_vectorSEXP);

    // ---- 1. the accelerated argument space; everything else stops with a message (nothing is silently ignored)
    if (n_starts != 1) stop("quilt-b200: n_gibbs_starts must be 1 (QUILT2 production value, functions.R:639)");
    if (!as<bool>(use_hapMatcherRSEXP)) stop("quilt-b200: only the hapMatcherR (raw) compressed panel is supported");
    static const char* const must_be_false[] = {"run_fb_subset", "return_gamma", "update_hapSum", "update_in_place", "use_small_eHapsCurrent_tc",
                                                "return_alpha", "return_extra", "return_p_store", "return_p1", "return_gibbs_block_output",
                                                "return_advanced_gibbs_block_output"};
    for (const char* f : must_be_false)
        if (flag(pl, f)) stop(std::string("quilt-b200: param_list$") + f + " = TRUE is not on the accelerated path");
    if (flag(pl, "pass_in_eMatRead_t") && !flag(pl, "make_eMatRead_t_rare_common"))
        stop("quilt-b200: a caller-supplied eMatRead_t is only accepted as the all-ones scratch of the rare/common call (rare_common.R:260)");
    if (!flag(pl, "use_eMatDH_special_symbols")) stop("quilt-b200: use_eMatDH_special_symbols = FALSE is not supported");
    if (!flag(pl, "haploid_gibbs_equal_weighting")) stop("quilt-b200: only equal weighting of sampling sweeps is supported");
    if (!flag(pl, "return_genProbs") || !flag(pl, "return_hapProbs")) stop("quilt-b200: return_genProbs and return_hapProbs must be TRUE");
    if (seed_vector.size() > 0 && seed_vector[0] > 0) stop("quilt-b200: seed_vector > 0 (reseeding inside the call) is not supported; production passes 0");
    if (as<bool>(do_block_resamplingSEXP)) stop("quilt-b200: do_block_resampling = TRUE is not supported (production passes FALSE, functions.R:2655)");
    if (as<bool>(generate_fb_snp_offsetsSEXP)) stop("quilt-b200: generate_fb_snp_offsets = TRUE is not supported");
    if (as<int>(run_fb_grid_offsetSEXP) != 0) stop("quilt-b200: run_fb_grid_offset must be 0");
    if (as<int>(snp_start_1_basedSEXP) != -1 || as<int>(snp_end_1_basedSEXP) != -1) stop("quilt-b200: snp_start_1_based / snp_end_1_based must be -1");
    if (as<int>(artificial_relabelSEXP) != -1) stop("quilt-b200: artificial_relabel must be -1");
    for (int i = 0; i < nSNPs; i++)
        if (grid[i] != i / 32) stop("quilt-b200: grid must be 32 SNPs per grid (grid[i] == i %/% 32, quilt-prepare-reference.R:376-380)");
    {
        const LogicalVector skip(skip_read_iterationSEXP);  // production: all FALSE unless small_ref_panel_skip_equally_likely_reads
        for (int i = 0; i < skip.size(); i++)
            if (skip[i]) stop("quilt-b200: skip_read_iteration = TRUE is not supported");
    }
    // (calculate_gamma_on_the_fly, pass_in_alphaBeta, verbose, suppressOutput, maxEmissionMatrixDifference, priorCurrent_m,
    //  alphaMatCurrent_tc, eHapsCurrent_tc, blocks_for_output, hapSum_tc, hapMatcher, rhb_t, prev_list_of_alphaBetaBlocks,
    //  i_snp_block_for_alpha_beta and the scratch matrices do not influence the results of the supported space.)

    // ---- 2. sampleReads -> CSR
    std::vector<int32_t> offsets(nReads + 1, 0), u, bq;
    for (int r = 0; r < nReads; r++) {
        const List rd = sampleReads[r];
        const IntegerVector bqv = as<IntegerVector>(rd[2]), uv = as<IntegerVector>(rd[3]);
        const int cnt = as<int>(rd[0]) + 1;  // J = count - 1
        for (int j = 0; j < cnt; j++) {
            u.push_back(uv[j]);
            bq.push_back(bqv[j]);
        }
        offsets[r + 1] = (int32_t)u.size();
    }
    const IntegerVector wif0(wif0SEXP);

    // ---- panel (the library keeps a device copy keyed by CONTENT, so call-local conversions below hit its cache)
    const RawMatrix hapMatcherR(hapMatcherRSEXP);
    const IntegerMatrix distinctHapsB(distinctHapsBSEXP);
    const NumericMatrix distinctHapsIE(distinctHapsIESEXP);
    const IntegerMatrix special(eMatDH_special_matrixSEXP), helper(eMatDH_special_matrix_helperSEXP);
    const bool rare_common = flag(pl, "make_eMatRead_t_rare_common");
    QuiltPanel panel;
    std::memset(&panel, 0, sizeof(panel));
    panel.K_full = hapMatcherR.nrow();
    panel.nGrids = hapMatcherR.ncol();
    panel.nSNPs = distinctHapsIE.ncol();
    panel.nMaxDH = distinctHapsB.nrow();
    panel.hapMatcherR = (const uint8_t*)RAW(hapMatcherRSEXP);
    panel.distinctHapsB = INTEGER(distinctHapsBSEXP);
    panel.distinctHapsIE = REAL(distinctHapsIESEXP);
    panel.eMatDH_special_matrix = INTEGER(eMatDH_special_matrixSEXP);
    panel.n_special = special.nrow();
    panel.eMatDH_special_matrix_helper = INTEGER(eMatDH_special_matrix_helperSEXP);
    panel.ref_error = as<double>(ref_errorSEXP);
    std::vector<uint8_t> is_common;
    std::vector<int64_t> rare_off;
    std::vector<int32_t> rare_snps;
    if (rare_common) {
        const LogicalVector sic(snp_is_commonSEXP);
        const List rph(rare_per_hap_infoSEXP);
        is_common.resize(sic.size());
        for (int i = 0; i < sic.size(); i++) is_common[i] = sic[i] ? 1 : 0;
        rare_off.assign(rph.size() + 1, 0);
        for (int k = 0; k < rph.size(); k++) {
            const IntegerVector v = as<IntegerVector>(rph[k]);
            for (int j = 0; j < v.size(); j++) rare_snps.push_back(v[j]);
            rare_off[k + 1] = (int64_t)rare_snps.size();
        }
        panel.nSNPs_all = sic.size();
        panel.snp_is_common = is_common.data();
        panel.common_snp_index = INTEGER(common_snp_indexSEXP);
        panel.rare_hap_offsets = rare_off.data();
        panel.rare_hap_snps = rare_snps.data();
    }

    // ---- flags
    uint32_t flags = 0;
    if (flag(pl, "sample_is_diploid")) flags |= QUILT_F_SAMPLE_IS_DIPLOID;
    if (flag(pl, "gibbs_initialize_iteratively")) flags |= QUILT_F_GIBBS_INITIALIZE_ITERATIVELY;
    if (flag(pl, "gibbs_initialize_at_first_read")) flags |= QUILT_F_GIBBS_INITIALIZE_AT_FIRST_READ;
    if (flag(pl, "perform_block_gibbs")) flags |= QUILT_F_PERFORM_BLOCK_GIBBS;
    if (flag(pl, "do_shard_block_gibbs")) flags |= QUILT_F_DO_SHARD_BLOCK_GIBBS;
    if (flag(pl, "shard_check_every_pair")) flags |= QUILT_F_SHARD_CHECK_EVERY_PAIR;
    if (flag(pl, "disable_read_category_usage")) flags |= QUILT_F_DISABLE_READ_CATEGORY_USAGE;
    if (flag(pl, "force_reset_read_category_zero")) flags |= QUILT_F_FORCE_RESET_READ_CATEGORY_0;
    if (rare_common) flags |= QUILT_F_MAKE_EMATREAD_RARE_COMMON;
    if (flag(pl, "rescale_eMatRead_t")) flags |= QUILT_F_RESCALE_EMATREAD;
    if (flag(pl, "record_read_set")) flags |= QUILT_F_RECORD_READ_SET;
    if (flag(pl, "use_smooth_cm_in_block_gibbs")) flags |= QUILT_F_USE_SMOOTH_CM_IN_BLOCK_GIBBS;
    const bool diploid = (flags & QUILT_F_SAMPLE_IS_DIPLOID) != 0;
    const bool shard = flag(pl, "do_shard_block_gibbs");
    if (shard && !diploid) stop("quilt-b200: the shard pass is diploid-only (functions.R:2552-2556)");
    if (shard && !flag(pl, "shard_check_every_pair")) stop("quilt-b200: shard pass without shard_check_every_pair is not supported (production: TRUE, quilt.R:178)");

    // ---- 3. R's generator, in the reference's order
    const NumericVector runif_reads = Rcpp::runif(nReads * n_full);                                  // gibbs-nipt.cpp:2845
    int first_read = 0;
    if (!flag(pl, "gibbs_initialize_at_first_read") && nReads > 0) first_read = Rcpp::sample(nReads, 1)(0) - 1;   // :2848
    std::vector<int32_t> H0(nReads, 1);
    if (flag(pl, "use_starting_read_labels")) {
        // double_list_of_starting_read_labels[[1]][[1]] (gibbs-nipt.cpp:2853-2861)
        const List outer(double_list_of_starting_read_labelsSEXP);
        const List inner = outer[0];
        const IntegerVector h = as<IntegerVector>(inner[0]);
        if (h.size() != nReads) stop("quilt-b200: starting read labels have the wrong length");
        for (int r = 0; r < nReads; r++) H0[r] = h[r];
    } else {
        // random_gibbs_nipt_read_labels (gibbs-nipt.cpp:1961-1978)
        if (diploid) stop("quilt-b200: use_starting_read_labels = FALSE draws label 3 for diploid samples in the reference; not supported");
        const NumericVector x = Rcpp::runif(nReads);
        for (int r = 0; r < nReads; r++) H0[r] = (x[r] < 0.5) ? 1 : (((0.5 <= x[r]) & (x[r] < (0.5 + ff / 2))) ? 2 : 3);
    }
    const IntegerVector block_its(block_gibbs_iterationsSEXP);
    std::vector<int32_t> bits;
    for (int i = 0; i < block_its.size(); i++)
        if (block_its[i] >= 0 && block_its[i] < n_full) bits.push_back(block_its[i]);
    const int n_ep = flag(pl, "perform_block_gibbs") ? (int)bits.size() : 0;
    // everything after this point is data-dependent: mark the generator, draw the longest stream the call can need
    const RngMark mark = rng_mark();
    const size_t per_ep = 8 * (size_t)nReads + (diploid ? 0 : (size_t)nReads) + ((shard && nGrids > 1) ? (size_t)(nGrids - 1) : 0);
    std::vector<double> stream((size_t)n_ep * per_ep);
    for (size_t i = 0; i < stream.size(); i++) stream[i] = unif_rand();

    // ---- 4. the call
    QuiltGibbsArgs a;
    std::memset(&a, 0, sizeof(a));
    a.panel = &panel;
    a.reads.nReads = nReads;
    a.reads.offsets = offsets.data();
    a.reads.u = u.data();
    a.reads.bq = bq.data();
    a.reads.wif0 = INTEGER(wif0SEXP);
    const IntegerVector which(which_haps_to_useSEXP);
    a.K = which.size();
    a.which_haps_to_use = INTEGER(which_haps_to_useSEXP);
    a.nGrids = nGrids;
    a.nSNPs = nSNPs;
    a.transMatRate_tc_H = REAL(transMatRate_tc_HSEXP);
    a.L_grid = INTEGER(L_gridSEXP);
    a.smooth_cm = REAL(smooth_cmSEXP);
    a.ff = ff;
    a.n_gibbs_burn_in_its = n_burn;
    a.n_gibbs_sample_its = n_sample;
    a.block_gibbs_iterations = bits.data();
    a.n_block_gibbs_iterations = n_ep;
    a.H0 = H0.data();
    a.first_read_for_gibbs_initialization = first_read;
    a.runif_reads = REAL(runif_reads);
    a.unif_stream = stream.empty() ? nullptr : stream.data();
    a.n_unif_stream = (int64_t)stream.size();
    a.maxDifferenceBetweenReads = as<double>(maxDifferenceBetweenReadsSEXP);
    a.Jmax = as<int>(Jmax_localSEXP);
    a.class_sum_cutoff = as<double>(class_sum_cutoffSEXP);
    a.shuffle_bin_radius = as<int>(shuffle_bin_radiusSEXP);
    a.block_gibbs_quantile_prob = as<double>(block_gibbs_quantile_probSEXP);
    a.flags = flags;

    NumericMatrix hapProbs_t(3, nSNPs), genProbsM_t(3, nSNPs), genProbsF_t(3, nSNPs);
    IntegerVector H(nReads), H_class(flag(pl, "record_read_set") ? nReads : 1);
    IntegerMatrix H_its(nReads, n_sample > 0 ? n_sample : 1);   // labels after each sampling sweep
    NumericMatrix per_it_likelihoods(n_sample == 0 ? 1 : n_full, 13);   // gibbs-nipt.cpp:2757-2767
    QuiltGibbsOut o;
    std::memset(&o, 0, sizeof(o));
    o.hapProbs_t = REAL(hapProbs_t);
    o.genProbsM_t = REAL(genProbsM_t);
    o.genProbsF_t = REAL(genProbsF_t);
    o.H = INTEGER(H);
    o.H_class = flag(pl, "record_read_set") ? INTEGER(H_class) : nullptr;
    o.per_it_likelihoods = REAL(per_it_likelihoods);
    o.H_sample_its = n_sample > 0 ? INTEGER(H_its) : nullptr;
    const int rc = QUILT_SHIM_BACKEND(&a, &o);
    if (rc != QUILT_OK) stop(std::string("quilt-b200: the Gibbs call failed: ") + QUILT_SHIM_LAST_ERROR());

    // leave R's generator where the reference would: rewind to the mark, advance by what the reference would have drawn
    if ((size_t)o.n_unif_consumed != stream.size()) {
        rng_rewind(mark);
        for (int64_t i = 0; i < o.n_unif_consumed; i++) (void)unif_rand();
    }

    // ---- 5. the reference's named list (gibbs-nipt.cpp:3217-3306)
    List to_return;
    if (o.underflow_problem) {
        to_return.push_back(true, "underflow_problem");  // gibbs-nipt.cpp:2963-2966
        return to_return;
    }
    to_return.push_back(false, "underflow_problem");
    to_return.push_back(genProbsM_t, "genProbsM_t");
    to_return.push_back(genProbsF_t, "genProbsF_t");
    to_return.push_back(hapProbs_t, "hapProbs_t");
    to_return.push_back(H, "H");
    List ending;   // double_list_of_ending_read_labels[[s]]$list_of_ending_read_labels: one "H" per sampling sweep (:3104, :3199)
    {
        List inner;
        for (int i = 0; i < n_sample; i++) {
            IntegerVector Hi(nReads);
            for (int r = 0; r < nReads; r++) Hi[r] = H_its(r, i);
            inner.push_back(Hi, "H");
        }
        if (n_full != 0) ending.push_back(inner, "list_of_ending_read_labels");
    }
    to_return.push_back(ending, "double_list_of_ending_read_labels");
    // names: gibbs-nipt.cpp:2768
    colnames(per_it_likelihoods) = CharacterVector::create("s", "i_samp", "i_it", "i_result_it", "p_O1_given_H1_L", "p_O2_given_H2_L",
                                                           "p_O3_given_H3_L", "p_O_given_H_L", "p_H_given_L", "p_O_H_given_L_up_to_C",
                                                           "p_set_H_given_L", "relabel", "p_H_class_given_L");
    to_return.push_back(per_it_likelihoods, "per_it_likelihoods");
    if (flag(pl, "record_read_set")) to_return.push_back(H_class, "H_class");
    return to_return;
    END_RCPP
}
