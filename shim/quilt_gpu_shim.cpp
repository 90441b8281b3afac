// quilt_gpu_shim.cpp — Rcpp translation unit that keeps the reference's .Call entry point for the Gibbs hot path
// and forwards it to the C ABI of libquiltgpu.so (include/quilt_b200.h).
//
// Built only where R + Rcpp exist (NOT in the build image of this repository: no R, no Rcpp headers; see
// INTEGRATION.md for the two-line Makevars change).  It replaces, symbol for symbol,
//     RcppExport SEXP _QUILT_rcpp_forwardBackwardGibbsNIPT(SEXP x 63)      QUILT/src/RcppExports.cpp:966-1038
// i.e. the glue of rcpp_forwardBackwardGibbsNIPT (QUILT/src/gibbs-nipt.cpp:2395-3307), so
// QUILT/R/RcppExports.R:215-217 and the production caller QUILT/R/functions.R:2614-2678 stay untouched.
//
// What the shim does, in the reference's order:
//   1. flattens sampleReads (R list of list(J, wif, bq, u), gibbs-small.cpp:149-152) into CSR arrays;
//   2. reads the 33 logicals of param_list (gibbs-nipt.cpp:2505-2536) into the flag word;
//   3. draws every uniform the reference draws from R's RNG inside the call, in the reference's order
//      (SURVEY.md §8b "RNG"): runif(nReads * n_full_its) (:2845), sample(nReads, 1) (:2848, only when
//      !gibbs_initialize_at_first_read), then per block-Gibbs episode six runif(nReads) rows + runif(nReads)
//      + runif(nReads) (:3013-3018; only the second-to-last is consumed) and, for diploid samples with the
//      shard pass, runif(nGrids - 1) (gibbs-nipt-block.cpp:2054);
//   4. calls quilt_gpu_gibbs (host pointers in / out, no R types);
//   5. rebuilds the named return list of gibbs-nipt.cpp:3217-3306 for the production param_list.
// The scratch matrices R passes in (alphaHat_t*, betaHat_t*, eMatGrid_t*, gamma*_t_local, eMatRead_t) are ignored:
// the device owns that state (R never reads them back in production, functions.R:2594 / quilt.R:731-762).
//
// Unsupported argument combinations (n_gibbs_starts > 1, run_fb_subset, return_gamma, update_hapSum, dense rhb_t
// panels, NIPT until the three-haplotype kernels land) are rejected with Rcpp::stop — never silently computed on
// the CPU.
#include <Rcpp.h>

#include <cstring>
#include <vector>

#include "../include/quilt_b200.h"

using namespace Rcpp;

namespace {

inline bool flag(const List& pl, const char* name) { return as<bool>(pl[name]); }

// Rcpp::sample(n, 1)(0): R's R_unif_index on the current stream (Rcpp sugar sample.h, non-replacement, size 1)
inline int sample_one(int n) { return Rcpp::sample(n, 1)(0); }

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
    const int nGrids = tm_dim[1] + 1;
    const IntegerVector grid(gridSEXP);
    const int nSNPs = grid.size();
    const int n_starts = as<int>(n_gibbs_startsSEXP);
    const int n_sample = as<int>(n_gibbs_sample_itsSEXP), n_burn = as<int>(n_gibbs_burn_in_itsSEXP);
    const int n_full = n_sample + n_burn;
    const double ff = as<double>(ffSEXP);
    const IntegerVector seed_vector(seed_vectorSEXP);

    if (n_starts != 1) stop("quilt-b200: n_gibbs_starts must be 1 (QUILT2 production value, functions.R:639)");
    if (!as<bool>(use_hapMatcherRSEXP)) stop("quilt-b200: only the hapMatcherR (raw) compressed panel is supported");
    if (flag(pl, "run_fb_subset") || flag(pl, "return_gamma") || flag(pl, "update_hapSum") || flag(pl, "pass_in_eMatRead_t"))
        stop("quilt-b200: run_fb_subset / return_gamma / update_hapSum / pass_in_eMatRead_t are not on the accelerated path");
    if (!flag(pl, "use_eMatDH_special_symbols")) stop("quilt-b200: use_eMatDH_special_symbols = FALSE is not supported");
    if (!flag(pl, "haploid_gibbs_equal_weighting")) stop("quilt-b200: only equal weighting of sampling sweeps is supported");
    if (seed_vector[0] > 0) stop("quilt-b200: seed_vector > 0 (reseeding inside the call) is not supported; production passes 0");

    // ---- 1. sampleReads -> CSR
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

    // ---- panel (pointers into R memory; the library uploads it once and caches it by pointer + shape)
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

    // ---- 2. flags
    uint32_t flags = 0;
    if (flag(pl, "sample_is_diploid")) flags |= QUILT_F_SAMPLE_IS_DIPLOID;
    if (flag(pl, "gibbs_initialize_iteratively")) flags |= QUILT_F_GIBBS_INITIALIZE_ITERATIVELY;
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

    // ---- starting labels: double_list_of_starting_read_labels[[1]][[1]] when use_starting_read_labels (gibbs-nipt.cpp:2853-2861)
    std::vector<int32_t> H0(nReads, 1);
    if (flag(pl, "use_starting_read_labels")) {
        const List outer(double_list_of_starting_read_labelsSEXP);
        const List inner = outer[0];
        const IntegerVector h = as<IntegerVector>(inner[0]);
        for (int r = 0; r < nReads; r++) H0[r] = h[r];
    } else {
        stop("quilt-b200: use_starting_read_labels = FALSE (labels drawn inside the call) is not supported; production passes TRUE");
    }

    // ---- 3. RNG, in the reference's order
    const NumericVector runif_reads = Rcpp::runif(nReads * n_full);
    int first_read = 0;
    if (!flag(pl, "gibbs_initialize_at_first_read") && nReads > 0) first_read = sample_one(nReads) - 1;
    const IntegerVector block_its(block_gibbs_iterationsSEXP);
    std::vector<int32_t> bits;
    for (int i = 0; i < block_its.size(); i++)
        if (block_its[i] >= 0 && block_its[i] < n_full) bits.push_back(block_its[i]);
    const int n_ep = flag(pl, "perform_block_gibbs") ? (int)bits.size() : 0;
    std::vector<double> runif_block((size_t)std::max(n_ep, 1) * nReads), runif_shard((size_t)std::max(n_ep, 1) * std::max(nGrids - 1, 1));
    const bool shard = diploid && flag(pl, "do_shard_block_gibbs");
    for (int e = 0; e < n_ep; e++) {
        for (int j = 0; j < 6; j++) (void)Rcpp::runif(nReads);  // runif_proposed: drawn, never consumed (block_approach 6)
        const NumericVector rb = Rcpp::runif(nReads);
        std::copy(rb.begin(), rb.end(), runif_block.begin() + (size_t)e * nReads);
        (void)Rcpp::runif(nReads);  // runif_total: drawn, never consumed
        if (shard && nGrids > 1) {
            const NumericVector rs = Rcpp::runif(nGrids - 1);
            std::copy(rs.begin(), rs.end(), runif_shard.begin() + (size_t)e * (nGrids - 1));
        }
    }
    // NB: an underflow early-return (gibbs-nipt.cpp:2963-2966) stops the reference's draws at the failing sweep; the
    // retry loop (functions.R:2704-2714) then continues from a different stream position than this shim would.
    // That only matters for bit-replay after an underflow; the retry itself is statistically equivalent.

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
    a.runif_block = runif_block.data();
    a.runif_shard = runif_shard.data();
    a.runif_H_class = nullptr;
    a.maxDifferenceBetweenReads = as<double>(maxDifferenceBetweenReadsSEXP);
    a.Jmax = as<int>(Jmax_localSEXP);
    a.class_sum_cutoff = as<double>(class_sum_cutoffSEXP);
    a.shuffle_bin_radius = as<int>(shuffle_bin_radiusSEXP);
    a.block_gibbs_quantile_prob = as<double>(block_gibbs_quantile_probSEXP);
    a.flags = flags;

    NumericMatrix hapProbs_t(3, nSNPs), genProbsM_t(3, nSNPs), genProbsF_t(3, nSNPs);
    IntegerVector H(nReads), H_class(nReads);
    NumericMatrix per_it_likelihoods(n_full, 13);
    QuiltGibbsOut o;
    std::memset(&o, 0, sizeof(o));
    o.hapProbs_t = REAL(hapProbs_t);
    o.genProbsM_t = REAL(genProbsM_t);
    o.genProbsF_t = REAL(genProbsF_t);
    o.H = INTEGER(H);
    o.H_class = INTEGER(H_class);
    o.per_it_likelihoods = REAL(per_it_likelihoods);
    const int rc = quilt_gpu_gibbs(&a, &o);
    if (rc != QUILT_OK) stop(std::string("quilt-b200: quilt_gpu_gibbs failed: ") + quilt_gpu_last_error());

    // ---- 5. the reference's named list (gibbs-nipt.cpp:3217-3306)
    List to_return;
    if (o.underflow_problem) {
        to_return.push_back(true, "underflow_problem");  // gibbs-nipt.cpp:2963-2966
        return to_return;
    }
    to_return.push_back(false, "underflow_problem");
    if (flag(pl, "return_genProbs")) {
        to_return.push_back(genProbsM_t, "genProbsM_t");
        to_return.push_back(genProbsF_t, "genProbsF_t");
    }
    if (flag(pl, "return_hapProbs")) to_return.push_back(hapProbs_t, "hapProbs_t");
    to_return.push_back(H, "H");
    List ending(1);
    {
        List inner(1);
        inner[0] = clone(H);
        ending[0] = inner;
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
