// marshal.h — flat C-ABI arguments (include/quilt_b200.h) -> the R objects the reference's entry point takes.
// TEST INFRASTRUCTURE (used by ref_cabi.cpp and by the shim harness); follows the production caller
// (QUILT/R/functions.R:2566-2678, QUILT/R/quilt.R:729-762, QUILT/R/rare_common.R:313-391).
#ifndef REFSHIM_MARSHAL_H
#define REFSHIM_MARSHAL_H
#include <RcppArmadillo.h>

#include "../../include/quilt_b200.h"

namespace refmarshal {

template <int RT, class T>
Rcpp::Vector<RT> foreign_vector(const T* p, size_t n) {
    return Rcpp::Vector<RT>(refshim::wrap_foreign(RT, const_cast<T*>(p), n));
}
template <int RT, class T>
Rcpp::Matrix<RT> foreign_matrix(const T* p, int nr, int nc) {
    SEXP s = refshim::wrap_foreign(RT, const_cast<T*>(p), (size_t)nr * nc);
    s->dim = {nr, nc};
    return Rcpp::Matrix<RT>(s);
}

// sampleReads as the R list of list(J, wif, bq, u) (test-drivers.R:222-227; bq / u are one-column integer matrices)
Rcpp::List make_sampleReads(const QuiltReads& r) {
    Rcpp::List out(r.nReads);
    for (int i = 0; i < r.nReads; ++i) {
        const int a = r.offsets[i], n = r.offsets[i + 1] - a;
        Rcpp::IntegerMatrix bq(n, 1), u(n, 1);
        for (int j = 0; j < n; ++j) { bq(j, 0) = r.bq[a + j]; u(j, 0) = r.u[a + j]; }
        out[i] = Rcpp::List::create(n - 1, (int)r.wif0[i], bq, u);
    }
    return out;
}

struct PanelObjects {
    arma::imat hapMatcher;
    Rcpp::RawMatrix hapMatcherR;
    arma::imat distinctHapsB;
    arma::mat distinctHapsIE;
    Rcpp::IntegerMatrix special_helper, special_matrix;
    arma::imat rhb_t;
    Rcpp::List rare_per_hap_info, rare_per_snp_info;
    Rcpp::IntegerVector common_snp_index;
    Rcpp::LogicalVector snp_is_common;
    PanelObjects(const QuiltPanel* p, bool rare_common, int K, const int32_t* which)
        : hapMatcher(1, 1),
          hapMatcherR(foreign_matrix<Rcpp::RAWSXP>(p->hapMatcherR, p->K_full, p->nGrids)),
          distinctHapsB(const_cast<int*>(p->distinctHapsB), (arma::uword)p->nMaxDH, (arma::uword)p->nGrids, false, true),
          distinctHapsIE(const_cast<double*>(p->distinctHapsIE), (arma::uword)p->nMaxDH, (arma::uword)p->nSNPs, false, true),
          special_helper(foreign_matrix<Rcpp::INTSXP>(p->eMatDH_special_matrix_helper, p->nGrids, 2)),
          special_matrix(foreign_matrix<Rcpp::INTSXP>(p->eMatDH_special_matrix, p->n_special, 2)),
          rhb_t(1, 1) {
        if (rare_common) {
            // rare_per_hap_info: list[K_full] of 1-based all-SNP indices; rare_per_snp_info: list[nSNPs_all] of
            // c(-1, k...) with k 1-based WITHIN which_haps_to_use, appended in k order (rare_common.R:313-322)
            rare_per_hap_info = Rcpp::List(p->K_full);
            for (int h = 0; h < p->K_full; ++h) {
                const int64_t a = p->rare_hap_offsets[h], b = p->rare_hap_offsets[h + 1];
                Rcpp::IntegerVector v((int)(b - a));
                for (int64_t j = a; j < b; ++j) v[j - a] = p->rare_hap_snps[j];
                rare_per_hap_info[h] = v;
            }
            std::vector<std::vector<int> > per_snp((size_t)p->nSNPs_all, std::vector<int>(1, -1));
            for (int k = 0; k < K; ++k) {
                const int h = which[k] - 1;
                for (int64_t j = p->rare_hap_offsets[h]; j < p->rare_hap_offsets[h + 1]; ++j) per_snp[(size_t)p->rare_hap_snps[j] - 1].push_back(k + 1);
            }
            rare_per_snp_info = Rcpp::List(p->nSNPs_all);
            for (int s = 0; s < p->nSNPs_all; ++s) rare_per_snp_info[s] = Rcpp::wrap(per_snp[(size_t)s]);
            common_snp_index = foreign_vector<Rcpp::INTSXP>(p->common_snp_index, (size_t)p->nSNPs_all);
            snp_is_common = Rcpp::LogicalVector(p->nSNPs_all);
            for (int s = 0; s < p->nSNPs_all; ++s) snp_is_common[s] = p->snp_is_common[s] ? 1 : 0;
        } else {
            // the reference's default arguments (gibbs-nipt.cpp:2455-2458)
            rare_per_hap_info = Rcpp::List::create(0);
            rare_per_snp_info = Rcpp::List::create(0);
            common_snp_index = Rcpp::IntegerVector::create(0);
            snp_is_common = Rcpp::LogicalVector::create(0);
        }
    }
};

Rcpp::IntegerVector make_grid(int nSNPs) {   // "grid32": SNP -> grid (quilt-prepare-reference.R:376-380)
    Rcpp::IntegerVector g(nSNPs);
    for (int i = 0; i < nSNPs; ++i) g[i] = i / 32;
    return g;
}


// every argument of rcpp_forwardBackwardGibbsNIPT for one flat call, built the way R builds them
struct CallObjects {
    int K, nGrids, nReads, nSNPs, n_full;
    bool diploid, rare_common;
    Rcpp::List sampleReads;
    PanelObjects P;
    arma::mat eMatRead_t, priorCurrent_m, blocks_for_output;
    arma::cube alphaMatCurrent_tc, eHapsCurrent_tc, transMatRate_tc_H, hapSum_tc;
    arma::mat alphaHat_t1, betaHat_t1, eMatGrid_t1, alphaHat_t2, betaHat_t2, eMatGrid_t2, alphaHat_t3, betaHat_t3, eMatGrid_t3;
    arma::mat gammaMT_t_local, gammaMU_t_local, gammaP_t_local;
    Rcpp::IntegerVector which_haps_to_use, wif0, L_grid, grid, block_its;
    Rcpp::LogicalVector grid_has_read, skip_read_iteration;
    Rcpp::NumericVector smooth_cm;
    Rcpp::List param_list, double_list_of_starting_read_labels;

    explicit CallObjects(const QuiltGibbsArgs* a)
        : K(a->K), nGrids(a->nGrids), nReads(a->reads.nReads), nSNPs(a->nSNPs), n_full(a->n_gibbs_burn_in_its + a->n_gibbs_sample_its),
          diploid((a->flags & QUILT_F_SAMPLE_IS_DIPLOID) != 0), rare_common((a->flags & QUILT_F_MAKE_EMATREAD_RARE_COMMON) != 0),
          sampleReads(make_sampleReads(a->reads)), P(a->panel, rare_common, a->K, a->which_haps_to_use),
          // scratch owned by R in production (quilt.R:729-762); hap 3 is 1 x 1 for diploid methods; eMatRead_t is all ones for
          // the rare/common call (rare_common.R:260) and a 1 x 1 dummy otherwise (functions.R:2545-2550)
          eMatRead_t(rare_common ? arma::mat(a->K, a->reads.nReads, arma::fill::ones) : arma::mat(1, 1)),
          priorCurrent_m(a->K, 1), blocks_for_output(1, 1), alphaMatCurrent_tc(a->K, a->nGrids - 1, 1), eHapsCurrent_tc(1, 1, 1),
          transMatRate_tc_H(const_cast<double*>(a->transMatRate_tc_H), 2, a->nGrids - 1, 1, true), hapSum_tc(1, 1, 1),
          alphaHat_t1(a->K, a->nGrids), betaHat_t1(a->K, a->nGrids), eMatGrid_t1(a->K, a->nGrids), alphaHat_t2(a->K, a->nGrids),
          betaHat_t2(a->K, a->nGrids), eMatGrid_t2(a->K, a->nGrids), alphaHat_t3(diploid ? 1 : a->K, diploid ? 1 : a->nGrids),
          betaHat_t3(diploid ? 1 : a->K, diploid ? 1 : a->nGrids), eMatGrid_t3(diploid ? 1 : a->K, diploid ? 1 : a->nGrids),
          gammaMT_t_local(1, 1), gammaMU_t_local(1, 1), gammaP_t_local(1, 1) {
        priorCurrent_m.fill(1 / double(K));
        alphaMatCurrent_tc.fill(1 / double(K));
        which_haps_to_use = foreign_vector<Rcpp::INTSXP>(a->which_haps_to_use, (size_t)K);
        wif0 = foreign_vector<Rcpp::INTSXP>(a->reads.wif0, (size_t)nReads);
        grid_has_read = Rcpp::LogicalVector(nGrids);   // functions.R:314-316
        for (int r = 0; r < nReads; ++r) grid_has_read[a->reads.wif0[r]] = 1;
        L_grid = foreign_vector<Rcpp::INTSXP>(a->L_grid, (size_t)nGrids);
        smooth_cm = foreign_vector<Rcpp::REALSXP>(a->smooth_cm, (size_t)(nGrids - 1));
        skip_read_iteration = Rcpp::LogicalVector(n_full);
        grid = make_grid(nSNPs);
        using Rcpp::Named;
        param_list = Rcpp::List::create(   // functions.R:2566-2599
            Named("return_alpha") = (a->flags & QUILT_F_RETURN_ALPHA) != 0, Named("return_extra") = (a->flags & QUILT_F_RETURN_EXTRA) != 0,
            Named("return_genProbs") = true, Named("return_gamma") = false, Named("return_hapProbs") = true, Named("return_p_store") = false,
            Named("return_p1") = false, Named("return_gibbs_block_output") = false,
            Named("return_advanced_gibbs_block_output") = false, Named("use_starting_read_labels") = true,
            Named("verbose") = false, Named("run_fb_subset") = false, Named("haploid_gibbs_equal_weighting") = true,
            Named("gibbs_initialize_iteratively") = (a->flags & QUILT_F_GIBBS_INITIALIZE_ITERATIVELY) != 0,
            Named("gibbs_initialize_at_first_read") = (a->flags & QUILT_F_GIBBS_INITIALIZE_AT_FIRST_READ) != 0,
            Named("use_smooth_cm_in_block_gibbs") = (a->flags & QUILT_F_USE_SMOOTH_CM_IN_BLOCK_GIBBS) != 0,
            Named("use_small_eHapsCurrent_tc") = false, Named("sample_is_diploid") = diploid, Named("update_in_place") = false,
            Named("do_shard_block_gibbs") = (a->flags & QUILT_F_DO_SHARD_BLOCK_GIBBS) != 0,
            Named("shard_check_every_pair") = (a->flags & QUILT_F_SHARD_CHECK_EVERY_PAIR) != 0,
            Named("force_reset_read_category_zero") = (a->flags & QUILT_F_FORCE_RESET_READ_CATEGORY_0) != 0,
            Named("disable_read_category_usage") = (a->flags & QUILT_F_DISABLE_READ_CATEGORY_USAGE) != 0,
            Named("calculate_gamma_on_the_fly") = true, Named("rescale_eMatRead_t") = (a->flags & QUILT_F_RESCALE_EMATREAD) != 0,
            Named("pass_in_eMatRead_t") = rare_common, Named("make_eMatRead_t_rare_common") = rare_common,
            Named("pass_in_alphaBeta") = true, Named("update_hapSum") = false,
            Named("record_read_set") = (a->flags & QUILT_F_RECORD_READ_SET) != 0,
            Named("perform_block_gibbs") = (a->flags & QUILT_F_PERFORM_BLOCK_GIBBS) != 0, Named("use_eMatDH_special_symbols") = true);
        Rcpp::IntegerVector H0(nReads);
        for (int r = 0; r < nReads; ++r) H0[r] = a->H0[r];
        double_list_of_starting_read_labels = Rcpp::List::create(Rcpp::List::create(H0));
        block_its = Rcpp::IntegerVector(a->n_block_gibbs_iterations);
        for (int i = 0; i < a->n_block_gibbs_iterations; ++i) block_its[i] = a->block_gibbs_iterations[i];
    }
};

}  // namespace refmarshal
#endif
