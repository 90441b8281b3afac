// shim_harness.cpp — executes shim/quilt_gpu_shim.cpp WITHOUT R (TEST INFRASTRUCTURE).
//
// One shared library holds (a) the Rcpp shim compiled against the stand-in headers with the CPU oracle as back end,
// (b) the reference's own object files and (c) this driver.  For one flat call (include/quilt_b200.h) it builds the 63 R
// objects the way the production caller does (marshal.h), installs a position-addressable random generator and runs
// EITHER the reference's rcpp_forwardBackwardGibbsNIPT OR the shim's _QUILT_rcpp_forwardBackwardGibbsNIPT on them.
// tests/test_shim_executes.py compares the two named lists field by field and the position the generator is left at —
// i.e. the shim's promise that R's random stream continues exactly where the reference would leave it (NIPT's
// data-dependent draws and the underflow early return included).
#include "marshal.h"

using namespace refmarshal;

Rcpp::List rcpp_forwardBackwardGibbsNIPT(
    const Rcpp::List& sampleReads, arma::mat& eMatRead_t, const arma::mat& priorCurrent_m, const arma::cube& alphaMatCurrent_tc,
    const arma::cube& eHapsCurrent_tc, const arma::cube& transMatRate_tc_H, const double ff, const arma::mat& blocks_for_output,
    arma::mat& alphaHat_t1, arma::mat& betaHat_t1, arma::mat& alphaHat_t2, arma::mat& betaHat_t2, arma::mat& alphaHat_t3,
    arma::mat& betaHat_t3, arma::mat& eMatGrid_t1, arma::mat& eMatGrid_t2, arma::mat& eMatGrid_t3, arma::mat& gammaMT_t_local,
    arma::mat& gammaMU_t_local, arma::mat& gammaP_t_local, arma::cube& hapSum_tc, arma::imat& hapMatcher,
    Rcpp::RawMatrix& hapMatcherR, bool use_hapMatcherR, arma::imat& distinctHapsB, arma::mat& distinctHapsIE,
    Rcpp::IntegerMatrix& eMatDH_special_matrix_helper, Rcpp::IntegerMatrix& eMatDH_special_matrix, const arma::imat& rhb_t,
    double ref_error, const Rcpp::IntegerVector& which_haps_to_use, Rcpp::IntegerVector& wif0, Rcpp::LogicalVector& grid_has_read,
    Rcpp::IntegerVector& L_grid, Rcpp::NumericVector& smooth_cm, Rcpp::List param_list, Rcpp::LogicalVector& skip_read_iteration,
    const int Jmax_local, const double maxDifferenceBetweenReads, const double maxEmissionMatrixDifference,
    const int run_fb_grid_offset, const Rcpp::IntegerVector& grid, int snp_start_1_based, int snp_end_1_based,
    const bool generate_fb_snp_offsets, const int suppressOutput, int n_gibbs_starts, const int n_gibbs_sample_its,
    const int n_gibbs_burn_in_its, const Rcpp::List& double_list_of_starting_read_labels, Rcpp::IntegerVector seed_vector,
    const Rcpp::List& prev_list_of_alphaBetaBlocks, const int i_snp_block_for_alpha_beta, const bool do_block_resampling,
    const int artificial_relabel, const double class_sum_cutoff, const int shuffle_bin_radius,
    const Rcpp::IntegerVector block_gibbs_iterations, const double block_gibbs_quantile_prob, const Rcpp::List& rare_per_hap_info,
    const Rcpp::IntegerVector& common_snp_index, const Rcpp::LogicalVector& snp_is_common, const Rcpp::List& rare_per_snp_info);

extern "C" SEXP _QUILT_rcpp_forwardBackwardGibbsNIPT(
    SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP,
    SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP,
    SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);

namespace {

// counter-based generator: value i of the stream is a hash of (seed, i), so "saving the state" is remembering i
class CounterRng : public refshim::RngSource {
    uint64_t seed;
    long long pos = 0;
    static uint64_t mix(uint64_t z) {
        z += 0x9e3779b97f4a7c15ull;
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        return z ^ (z >> 31);
    }
public:
    explicit CounterRng(uint64_t s) : seed(s) {}
    double unif_rand() override {
        const uint64_t z = mix(seed ^ mix((uint64_t)pos++));
        return ((double)(z >> 11) + 0.5) * (1.0 / 9007199254740992.0);   // in (0, 1), like R's fixed-up unif_rand()
    }
    long long save() override { return pos; }
    void restore(long long p) override { pos = p; }
    long long position() const { return pos; }
};

thread_local std::string g_err;

}  // namespace

extern "C" {

const char* shim_harness_last_error(void) { return g_err.c_str(); }

// via = 0: the reference; via = 1: the shim (oracle back end).  Returns 0, or 2 with the message in shim_harness_last_error.
int shim_harness_run(const QuiltGibbsArgs* a, uint64_t seed, int via, QuiltGibbsOut* o, int64_t* rng_position_after, int32_t* n_list_elements) {
    try {
        CallObjects C(a);
        CounterRng rng(seed);
        refshim::RngSource* prev = refshim::rng_slot();
        refshim::rng_slot() = &rng;
        Rcpp::List out;
        try {
            if (via == 0) {
                out = rcpp_forwardBackwardGibbsNIPT(
                    C.sampleReads, C.eMatRead_t, C.priorCurrent_m, C.alphaMatCurrent_tc, C.eHapsCurrent_tc, C.transMatRate_tc_H, a->ff,
                    C.blocks_for_output, C.alphaHat_t1, C.betaHat_t1, C.alphaHat_t2, C.betaHat_t2, C.alphaHat_t3, C.betaHat_t3, C.eMatGrid_t1,
                    C.eMatGrid_t2, C.eMatGrid_t3, C.gammaMT_t_local, C.gammaMU_t_local, C.gammaP_t_local, C.hapSum_tc, C.P.hapMatcher, C.P.hapMatcherR,
                    true, C.P.distinctHapsB, C.P.distinctHapsIE, C.P.special_helper, C.P.special_matrix, C.P.rhb_t, a->panel->ref_error,
                    C.which_haps_to_use, C.wif0, C.grid_has_read, C.L_grid, C.smooth_cm, C.param_list, C.skip_read_iteration, a->Jmax,
                    a->maxDifferenceBetweenReads, 1e10, 0, C.grid, -1, -1, false, 1, 1, a->n_gibbs_sample_its, a->n_gibbs_burn_in_its,
                    C.double_list_of_starting_read_labels, Rcpp::IntegerVector::create(0), Rcpp::List::create(1, 2), -1, false, -1,
                    a->class_sum_cutoff, a->shuffle_bin_radius, C.block_its, a->block_gibbs_quantile_prob, C.P.rare_per_hap_info,
                    C.P.common_snp_index, C.P.snp_is_common, C.P.rare_per_snp_info);
            } else {
                using Rcpp::wrap;
                out = Rcpp::List(_QUILT_rcpp_forwardBackwardGibbsNIPT(
                    wrap(C.sampleReads), wrap(C.eMatRead_t), wrap(C.priorCurrent_m), wrap(C.alphaMatCurrent_tc), wrap(C.eHapsCurrent_tc),
                    wrap(C.transMatRate_tc_H), wrap(a->ff), wrap(C.blocks_for_output), wrap(C.alphaHat_t1), wrap(C.betaHat_t1), wrap(C.alphaHat_t2),
                    wrap(C.betaHat_t2), wrap(C.alphaHat_t3), wrap(C.betaHat_t3), wrap(C.eMatGrid_t1), wrap(C.eMatGrid_t2), wrap(C.eMatGrid_t3),
                    wrap(C.gammaMT_t_local), wrap(C.gammaMU_t_local), wrap(C.gammaP_t_local), wrap(C.hapSum_tc), wrap(C.P.hapMatcher),
                    wrap(C.P.hapMatcherR), wrap(true), wrap(C.P.distinctHapsB), wrap(C.P.distinctHapsIE), wrap(C.P.special_helper),
                    wrap(C.P.special_matrix), wrap(C.P.rhb_t), wrap(a->panel->ref_error), wrap(C.which_haps_to_use), wrap(C.wif0),
                    wrap(C.grid_has_read), wrap(C.L_grid), wrap(C.smooth_cm), wrap(C.param_list), wrap(C.skip_read_iteration), wrap(a->Jmax),
                    wrap(a->maxDifferenceBetweenReads), wrap(1e10), wrap(0), wrap(C.grid), wrap(-1), wrap(-1), wrap(false), wrap(1), wrap(1),
                    wrap(a->n_gibbs_sample_its), wrap(a->n_gibbs_burn_in_its), wrap(C.double_list_of_starting_read_labels),
                    wrap(Rcpp::IntegerVector::create(0)), wrap(Rcpp::List::create(1, 2)), wrap(-1), wrap(false), wrap(-1), wrap(a->class_sum_cutoff),
                    wrap(a->shuffle_bin_radius), wrap(C.block_its), wrap(a->block_gibbs_quantile_prob), wrap(C.P.rare_per_hap_info),
                    wrap(C.P.common_snp_index), wrap(C.P.snp_is_common), wrap(C.P.rare_per_snp_info)));
            }
        } catch (...) {
            refshim::rng_slot() = prev;
            throw;
        }
        refshim::rng_slot() = prev;
        *rng_position_after = rng.position();
        *n_list_elements = out.size();
        o->underflow_problem = Rcpp::as<bool>(out["underflow_problem"]) ? 1 : 0;
        if (o->underflow_problem) return 0;
        const int nSNPs = a->nSNPs, nReads = a->reads.nReads;
        auto copy_mat = [&](const char* name, double* dst, size_t n) {
            Rcpp::NumericMatrix m = Rcpp::as<Rcpp::NumericMatrix>(out[name]);
            if ((size_t)m.size() != n) throw std::logic_error(std::string("unexpected size of ") + name);
            if (dst) std::memcpy(dst, m.begin(), sizeof(double) * n);
        };
        copy_mat("hapProbs_t", o->hapProbs_t, (size_t)3 * nSNPs);
        copy_mat("genProbsM_t", o->genProbsM_t, (size_t)3 * nSNPs);
        copy_mat("genProbsF_t", o->genProbsF_t, (size_t)3 * nSNPs);
        Rcpp::IntegerVector H = Rcpp::as<Rcpp::IntegerVector>(out["H"]);
        // R reads out$double_list_of_ending_read_labels[[1]][[1]] (functions.R:745): one label vector per sampling sweep
        Rcpp::List dl = Rcpp::as<Rcpp::List>(out["double_list_of_ending_read_labels"]);
        Rcpp::List inner = Rcpp::as<Rcpp::List>(dl[0]);
        if (inner.size() != a->n_gibbs_sample_its) throw std::logic_error("double_list_of_ending_read_labels[[1]] has the wrong length");
        for (int i = 0; i < a->n_gibbs_sample_its; ++i) {
            Rcpp::IntegerVector Hi = Rcpp::as<Rcpp::IntegerVector>(inner[i]);
            if (o->H_sample_its)
                for (int r = 0; r < nReads; ++r) o->H_sample_its[(size_t)i * nReads + r] = Hi[r];
            if (i == a->n_gibbs_sample_its - 1)
                for (int r = 0; r < nReads; ++r)
                    if (Hi[r] != H[r]) throw std::logic_error("the last ending read label vector differs from H");
        }
        for (int r = 0; r < nReads; ++r)
            if (o->H) o->H[r] = H[r];
        if (o->H_class && (a->flags & QUILT_F_RECORD_READ_SET)) {
            Rcpp::IntegerVector Hc = Rcpp::as<Rcpp::IntegerVector>(out["H_class"]);
            for (int r = 0; r < nReads; ++r) o->H_class[r] = Hc[r];
        }
        if (o->per_it_likelihoods) {
            Rcpp::NumericMatrix m = Rcpp::as<Rcpp::NumericMatrix>(out["per_it_likelihoods"]);
            if (m.sexp()->colnames.size() != 13) throw std::logic_error("per_it_likelihoods lost its column names");
            std::memcpy(o->per_it_likelihoods, m.begin(), sizeof(double) * (size_t)m.size());
        }
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 2;
    }
}

}  // extern "C"
