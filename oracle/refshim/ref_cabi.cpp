// ref_cabi.cpp — C-ABI driver around the UNMODIFIED reference sources (TEST INFRASTRUCTURE, not product code).
//
// Links against the object files compiled from /root/reference/QUILT/src/{copied-from-stitch, gibbs-small,
// gibbs-nipt, gibbs-nipt-block, reference-single}.cpp (built with the RcppArmadillo stand-in next to this file;
// recipe: oracle/refshim/Makefile, output oracle/_ref/libquiltref.so) and exposes them through the same flat
// structs as the CUDA library and the oracle (include/quilt_b200.h), prefix quilt_ref_:
//
//   quilt_ref_gibbs               -> rcpp_forwardBackwardGibbsNIPT            (gibbs-nipt.cpp:2395-3307)
//   quilt_ref_make_eMatRead_t     -> Rcpp_make_eMatRead_t_for_gibbs_using_objects (gibbs-small.cpp:116-265),
//                                    rare/common sibling (:275-465), rcpp_evaluate_read_variability (gibbs-nipt.cpp:338-382)
//   quilt_ref_forward_backward    -> rcpp_initialize_gibbs_forward_backward   (gibbs-nipt.cpp:450-487)
//   quilt_ref_unpack_panel        -> rcpp_int_expand over the panel words     (copied-from-stitch.cpp:50-69)
//
// The arguments are marshalled exactly as the production caller does (QUILT/R/functions.R:2566-2678,
// QUILT/R/quilt.R:729-762, QUILT/R/rare_common.R:313-391).  R's random stream is replaced by a script built from
// the uniforms the flat ABI carries, replayed in the reference's own draw order; a draw the script does not
// expect raises an error, so a stream misalignment cannot pass silently.
#include <RcppArmadillo.h>
#include <RcppEigen.h>

#include <deque>
#include <mutex>

#include "marshal.h"

// ---- declarations of the reference functions called below (the definitions live in the reference sources)
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

void Rcpp_make_eMatRead_t_for_gibbs_using_objects(
    arma::mat& eMatRead_t, const Rcpp::List& sampleReads, const arma::imat& hapMatcher, const Rcpp::RawMatrix hapMatcherR,
    const bool use_hapMatcherR, const Rcpp::IntegerVector& grid, const arma::imat& rhb_t, const arma::mat& distinctHapsIE,
    const Rcpp::IntegerMatrix& eMatDH_special_matrix_helper, const Rcpp::IntegerMatrix& eMatDH_special_matrix, const double ref_error,
    const Rcpp::IntegerVector& which_haps_to_use, const bool rescale_eMatRead_t, const int Jmax, const double maxDifferenceBetweenReads,
    const bool use_eMatDH_special_symbols);

void Rcpp_make_eMatRead_t_for_final_rare_common_gibbs_using_objects(
    arma::mat& eMatRead_t, const Rcpp::List& rare_per_hap_info, const Rcpp::IntegerVector& common_snp_index,
    const Rcpp::LogicalVector& snp_is_common, const Rcpp::List& sampleReads, const Rcpp::RawMatrix hapMatcherR,
    const Rcpp::IntegerVector& grid, const arma::mat& distinctHapsIE, const Rcpp::IntegerMatrix& eMatDH_special_matrix_helper,
    const Rcpp::IntegerMatrix& eMatDH_special_matrix, const double ref_error, const Rcpp::IntegerVector& which_haps_to_use,
    const bool rescale_eMatRead_t, const int Jmax, const double maxDifferenceBetweenReads, const Rcpp::List& rare_per_snp_info);

void rcpp_evaluate_read_variability(arma::mat& eMatRead_t, arma::ivec& number_of_non_1_reads, arma::imat& indices_of_non_1_reads,
                                    arma::ivec& read_category, int cutoff);

void rcpp_initialize_gibbs_forward_backward(const arma::cube& alphaMatCurrent_tc, const arma::cube& transMatRate_tc_H,
                                            const arma::mat& priorCurrent_m, int s, arma::mat& alphaHat_t, arma::mat& betaHat_t,
                                            arma::rowvec& c, arma::mat& eMatGrid_t, const bool run_fb_subset,
                                            const Rcpp::NumericVector alphaStart, const Rcpp::NumericVector betaEnd);

void Rcpp_haploid_dosage_versus_refs(
    const arma::mat& gl, arma::mat& arma_alphaHat_t, Eigen::Map<Eigen::MatrixXd> eigen_alphaHat_t, arma::mat& betaHat_t, arma::rowvec& c,
    arma::mat& gamma_t, arma::mat& gammaSmall_t, Rcpp::List& best_haps_stuff_list, Rcpp::NumericVector& dosage, const arma::mat& transMatRate_t,
    const arma::imat& rhb_t, double ref_error, const bool use_eMatDH, arma::imat& distinctHapsB, arma::mat& distinctHapsIE,
    Rcpp::IntegerMatrix& eMatDH_special_matrix_helper, Rcpp::IntegerMatrix& eMatDH_special_matrix, const bool use_eMatDH_special_symbols,
    arma::imat& hapMatcher, Rcpp::RawMatrix& hapMatcherR, bool use_hapMatcherR, Rcpp::IntegerVector& gammaSmall_cols_to_get,
    const Rcpp::IntegerVector& eMatDH_special_grid_which, const Rcpp::List& eMatDH_special_values_list, const int K_top_matches,
    const int suppressOutput, const double min_emission_prob_normalization_threshold, bool return_betaHat_t, bool return_dosage,
    bool return_gamma_t, bool return_gammaSmall_t, bool get_best_haps_from_thinned_sites, bool is_version_2, bool is_version_3,
    bool return_extra, bool always_normalize, bool use_eigen, const bool normalize_emissions);

Rcpp::IntegerVector rcpp_int_expand(arma::ivec& hapc, const int nSNPs);
int rcpp_simple_binary_matrix_search(int val, Rcpp::IntegerMatrix mat, int s1, int e1);

using namespace refmarshal;

namespace {

thread_local std::string g_err;

// ------------------------------------------------------------------ scripted random stream
// Segments are consumed in order.  RUNIF segments must be requested with exactly their length (Rcpp::runif(n));
// a SAMPLE_INT segment answers one Rcpp::sample(n, 1); WEIGHTED segments answer any number (0..len) of
// Rcpp::sample(x, 1, false, probs) draws and are closed by the next request of another kind.
class ScriptedRng : public refshim::RngSource {
public:
    enum Kind { RUNIF, RUNIF_DUMMY, SAMPLE_INT, WEIGHTED, STREAM };
    struct Seg { Kind kind; const double* p; long n; int value; long used; };
    std::deque<Seg> segs;
    long n_runif_calls = 0, n_weighted = 0;

    void push_runif(const double* p, long n) { segs.push_back(Seg{RUNIF, p, n, 0, 0}); }
    void push_dummy(long n) { segs.push_back(Seg{RUNIF_DUMMY, nullptr, n, 0, 0}); }
    void push_sample_int(int one_based_value) { segs.push_back(Seg{SAMPLE_INT, nullptr, 1, one_based_value, 0}); }
    void push_weighted(const double* p, long n) { segs.push_back(Seg{WEIGHTED, p, n, 0, 0}); }
    // a flat unif_rand() stream that answers every later request in whatever order the reference makes them
    void push_stream(const double* p, long n) { segs.push_back(Seg{STREAM, p, n, 0, 0}); }
    long stream_used() const { return (!segs.empty() && segs.front().kind == STREAM) ? segs.front().used : 0; }
    double* stream_take(long n) {
        Seg& s = segs.front();
        if (s.used + n > s.n) throw std::logic_error("ScriptedRng: the episode stream is too short");
        double* q = const_cast<double*>(s.p) + s.used;
        s.used += n;
        return q;
    }

    void skip_weighted() { while (!segs.empty() && segs.front().kind == WEIGHTED) segs.pop_front(); }
    double unif_rand() override { throw std::logic_error("ScriptedRng: bare unif_rand() is not scripted"); }
    void runif(int n, double* out) override {
        skip_weighted();
        if (!segs.empty() && segs.front().kind == STREAM) {
            std::memcpy(out, stream_take(n), sizeof(double) * (size_t)n);
            ++n_runif_calls;
            return;
        }
        if (segs.empty()) throw std::logic_error("ScriptedRng: runif(" + std::to_string(n) + ") requested past the end of the script");
        Seg s = segs.front();
        segs.pop_front();
        if ((s.kind != RUNIF && s.kind != RUNIF_DUMMY) || s.n != n)
            throw std::logic_error("ScriptedRng: runif(" + std::to_string(n) + ") does not match the scripted draw order (expected kind " +
                                   std::to_string((int)s.kind) + ", n = " + std::to_string(s.n) + ")");
        if (s.kind == RUNIF) std::memcpy(out, s.p, sizeof(double) * (size_t)n);
        else for (int i = 0; i < n; ++i) out[i] = 0.5;
        ++n_runif_calls;
    }
    int sample_int(int n) override {
        skip_weighted();
        if (segs.empty() || segs.front().kind != SAMPLE_INT) throw std::logic_error("ScriptedRng: sample(n, 1) does not match the scripted draw order");
        Seg s = segs.front();
        segs.pop_front();
        if (s.value < 1 || s.value > n) throw std::logic_error("ScriptedRng: scripted sample(n, 1) value out of range");
        return s.value;
    }
    double unif_rand_for_weighted_sample() override {
        if (!segs.empty() && segs.front().kind == STREAM) { ++n_weighted; return *stream_take(1); }
        if (segs.empty() || segs.front().kind != WEIGHTED) throw std::logic_error("ScriptedRng: weighted sample() does not match the scripted draw order");
        Seg& s = segs.front();
        if (s.used >= s.n) throw std::logic_error("ScriptedRng: more weighted sample() draws than scripted uniforms");
        ++n_weighted;
        return s.p[s.used++];
    }
};

struct RngInstall {
    refshim::RngSource* prev;
    explicit RngInstall(refshim::RngSource* r) : prev(refshim::rng_slot()) { refshim::rng_slot() = r; }
    ~RngInstall() { refshim::rng_slot() = prev; }
};

template <class F>
int guarded(F f) {
    try {
        return f();
    } catch (const std::exception& e) {
        g_err = e.what();
        std::fprintf(stderr, "quilt_ref: %s\n", e.what());
        return QUILT_ERR_BAD_ARG;
    }
}

}  // namespace

extern "C" {

const char* quilt_ref_last_error(void) { return g_err.c_str(); }
int quilt_ref_make_eMatRead_t(const QuiltGibbsArgs* a, double* eMatRead_out, int32_t* read_category);

int quilt_ref_gibbs(const QuiltGibbsArgs* a, QuiltGibbsOut* o) {
    if (!a || !o || !a->panel) return QUILT_ERR_BAD_ARG;
    return guarded([&]() -> int {
        const int K = a->K, nGrids = a->nGrids, nReads = a->reads.nReads, nSNPs = a->nSNPs;
        const bool diploid = (a->flags & QUILT_F_SAMPLE_IS_DIPLOID) != 0;
        const bool rare_common = (a->flags & QUILT_F_MAKE_EMATREAD_RARE_COMMON) != 0;
        const bool perform_block_gibbs = (a->flags & QUILT_F_PERFORM_BLOCK_GIBBS) != 0;
        const bool do_shard = (a->flags & QUILT_F_DO_SHARD_BLOCK_GIBBS) != 0;
        const bool shard_every_pair = (a->flags & QUILT_F_SHARD_CHECK_EVERY_PAIR) != 0;
        const bool return_alpha = (a->flags & QUILT_F_RETURN_ALPHA) != 0;
        const bool return_extra = (a->flags & QUILT_F_RETURN_EXTRA) != 0;
        const int n_full = a->n_gibbs_burn_in_its + a->n_gibbs_sample_its;

        CallObjects C(a);

        // R's random stream in the reference's draw order (SURVEY.md section 8b "RNG")
        ScriptedRng rng;
        rng.push_runif(a->runif_reads, (long)nReads * n_full);                       // gibbs-nipt.cpp:2845
        if (nReads > 0 && !(a->flags & QUILT_F_GIBBS_INITIALIZE_AT_FIRST_READ)) rng.push_sample_int(a->first_read_for_gibbs_initialization + 1);   // :2848
        if (a->unif_stream) {
            rng.push_stream(a->unif_stream, (long)a->n_unif_stream);
        } else if (perform_block_gibbs) {
            int episode = 0;
            for (int it = 0; it < n_full; ++it) {
                bool hit = false;
                for (int i = 0; i < a->n_block_gibbs_iterations; ++i) hit |= (a->block_gibbs_iterations[i] == it);
                if (!hit) continue;
                for (int j = 0; j < 6; ++j) rng.push_dummy(nReads);                   // runif_proposed, never consumed (:3013-3016)
                rng.push_runif(a->runif_block + (size_t)episode * nReads, nReads);    // :3017
                rng.push_dummy(nReads);                                               // runif_total (:3018)
                if (!diploid && a->runif_H_class)                                     // rcpp_sample_H_using_H_class, block.cpp:226-243
                    rng.push_weighted(a->runif_H_class + (size_t)episode * nReads, nReads);
                if (do_shard) {                                                       // block.cpp:2054: runif(n_blocks - 1)
                    if (!shard_every_pair) throw std::logic_error("quilt_ref_gibbs: shard pass without shard_check_every_pair needs a data-dependent runif length; not scripted");
                    rng.push_runif(a->runif_shard + (size_t)episode * (nGrids - 1), nGrids - 1);
                }
                ++episode;
            }
        }
        RngInstall install(&rng);

        Rcpp::List out = rcpp_forwardBackwardGibbsNIPT(
            C.sampleReads, C.eMatRead_t, C.priorCurrent_m, C.alphaMatCurrent_tc, C.eHapsCurrent_tc, C.transMatRate_tc_H, a->ff, C.blocks_for_output,
            C.alphaHat_t1, C.betaHat_t1, C.alphaHat_t2, C.betaHat_t2, C.alphaHat_t3, C.betaHat_t3, C.eMatGrid_t1, C.eMatGrid_t2, C.eMatGrid_t3,
            C.gammaMT_t_local, C.gammaMU_t_local, C.gammaP_t_local, C.hapSum_tc, C.P.hapMatcher, C.P.hapMatcherR, true, C.P.distinctHapsB,
            C.P.distinctHapsIE, C.P.special_helper, C.P.special_matrix, C.P.rhb_t, a->panel->ref_error, C.which_haps_to_use, C.wif0, C.grid_has_read,
            C.L_grid, C.smooth_cm, C.param_list, C.skip_read_iteration, a->Jmax, a->maxDifferenceBetweenReads, 1e10 /*maxEmissionMatrixDifference*/,
            0 /*run_fb_grid_offset*/, C.grid, -1, -1, false /*generate_fb_snp_offsets*/, 1 /*suppressOutput*/, 1 /*n_gibbs_starts*/,
            a->n_gibbs_sample_its, a->n_gibbs_burn_in_its, C.double_list_of_starting_read_labels, Rcpp::IntegerVector::create(0) /*seed_vector*/,
            Rcpp::List::create(1, 2) /*prev_list_of_alphaBetaBlocks*/, -1, false /*do_block_resampling*/, -1 /*artificial_relabel*/,
            a->class_sum_cutoff, a->shuffle_bin_radius, C.block_its, a->block_gibbs_quantile_prob, C.P.rare_per_hap_info, C.P.common_snp_index,
            C.P.snp_is_common, C.P.rare_per_snp_info);

        o->underflow_problem = Rcpp::as<bool>(out["underflow_problem"]) ? 1 : 0;
        o->n_unif_consumed = a->unif_stream ? rng.stream_used() : 0;
        o->underflow_iteration = o->underflow_problem ? -2 /* the reference does not say which sweep */ : -1;
        if (o->underflow_problem) return QUILT_OK;   // list(underflow_problem = TRUE) early return, gibbs-nipt.cpp:2963-2966
        rng.skip_weighted();
        if (!rng.segs.empty() && rng.segs.front().kind != ScriptedRng::STREAM)
            throw std::logic_error("quilt_ref_gibbs: the reference consumed fewer random draws than scripted");

        auto copy_mat = [&](const char* name, double* dst, size_t n) {
            if (!dst) return;
            Rcpp::NumericMatrix m = Rcpp::as<Rcpp::NumericMatrix>(out[name]);
            if ((size_t)m.size() != n) throw std::logic_error(std::string("quilt_ref_gibbs: unexpected size of ") + name);
            std::memcpy(dst, m.begin(), sizeof(double) * n);
        };
        copy_mat("hapProbs_t", o->hapProbs_t, (size_t)3 * nSNPs);
        copy_mat("genProbsM_t", o->genProbsM_t, (size_t)3 * nSNPs);
        copy_mat("genProbsF_t", o->genProbsF_t, (size_t)3 * nSNPs);
        if (o->H) {
            Rcpp::IntegerVector H = Rcpp::as<Rcpp::IntegerVector>(out["H"]);
            for (int r = 0; r < nReads; ++r) o->H[r] = H[r];
        }
        if (o->H_sample_its && a->n_gibbs_sample_its > 0) {
            // double_list_of_ending_read_labels[[1]][[i]]
            Rcpp::List inner = Rcpp::as<Rcpp::List>(Rcpp::as<Rcpp::List>(out["double_list_of_ending_read_labels"])[0]);
            if (inner.size() != a->n_gibbs_sample_its) throw std::logic_error("quilt_ref_gibbs: unexpected number of ending read label vectors");
            for (int i = 0; i < a->n_gibbs_sample_its; ++i) {
                Rcpp::IntegerVector Hi = Rcpp::as<Rcpp::IntegerVector>(inner[i]);
                for (int r = 0; r < nReads; ++r) o->H_sample_its[(size_t)i * nReads + r] = Hi[r];
            }
        }
        if (o->H_class && (a->flags & QUILT_F_RECORD_READ_SET)) {
            Rcpp::IntegerVector Hc = Rcpp::as<Rcpp::IntegerVector>(out["H_class"]);
            for (int r = 0; r < nReads; ++r) o->H_class[r] = Hc[r];
        }
        if (o->per_it_likelihoods) {
            Rcpp::NumericMatrix m = Rcpp::as<Rcpp::NumericMatrix>(out["per_it_likelihoods"]);
            std::memcpy(o->per_it_likelihoods, m.begin(), sizeof(double) * (size_t)m.size());
        }
        if (return_alpha) {
            arma::mat* al[3] = {&C.alphaHat_t1, &C.alphaHat_t2, &C.alphaHat_t3};
            arma::mat* be[3] = {&C.betaHat_t1, &C.betaHat_t2, &C.betaHat_t3};
            arma::mat* eg[3] = {&C.eMatGrid_t1, &C.eMatGrid_t2, &C.eMatGrid_t3};
            const char* cn[3] = {"c1", "c2", "c3"};
            for (int h = 0; h < (diploid ? 2 : 3); ++h) {
                if (o->alphaHat_t[h]) std::memcpy(o->alphaHat_t[h], al[h]->memptr(), sizeof(double) * (size_t)K * nGrids);
                if (o->betaHat_t[h]) std::memcpy(o->betaHat_t[h], be[h]->memptr(), sizeof(double) * (size_t)K * nGrids);
                if (o->eMatGrid_t[h]) std::memcpy(o->eMatGrid_t[h], eg[h]->memptr(), sizeof(double) * (size_t)K * nGrids);
                if (o->c[h]) {
                    Rcpp::NumericVector c = Rcpp::as<Rcpp::NumericVector>(out[cn[h]]);
                    std::memcpy(o->c[h], c.begin(), sizeof(double) * (size_t)nGrids);
                }
            }
        }
        if (return_extra && o->eMatRead_t) {
            Rcpp::NumericMatrix m = Rcpp::as<Rcpp::NumericMatrix>(out["eMatRead_t"]);
            std::memcpy(o->eMatRead_t, m.begin(), sizeof(double) * (size_t)K * nReads);
        }
        if (o->read_category) {
            // not part of the reference's return list: recomputed from the returned/updated eMatRead_t the way
            // the driver does (gibbs-nipt.cpp:2870-2890: evaluate, then the two overrides)
            std::vector<double> tmp((size_t)K * nReads);
            int rc = quilt_ref_make_eMatRead_t(a, tmp.data(), o->read_category);
            if (rc != QUILT_OK) return rc;
        }
        return QUILT_OK;
    });
}

int quilt_ref_make_eMatRead_t(const QuiltGibbsArgs* a, double* eMatRead_out, int32_t* read_category) {
    if (!a || !a->panel || !eMatRead_out) return QUILT_ERR_BAD_ARG;
    return guarded([&]() -> int {
        const int K = a->K, nReads = a->reads.nReads;
        const bool rare_common = (a->flags & QUILT_F_MAKE_EMATREAD_RARE_COMMON) != 0;
        const bool rescale = (a->flags & QUILT_F_RESCALE_EMATREAD) != 0;
        Rcpp::List sampleReads = make_sampleReads(a->reads);
        PanelObjects P(a->panel, rare_common, K, a->which_haps_to_use);
        Rcpp::IntegerVector which_haps_to_use = foreign_vector<Rcpp::INTSXP>(a->which_haps_to_use, (size_t)K);
        Rcpp::IntegerVector grid = make_grid(a->nSNPs);
        arma::mat eMatRead_t(eMatRead_out, (arma::uword)K, (arma::uword)nReads, false, true);
        eMatRead_t.fill(1.0);
        if (rare_common) {
            Rcpp_make_eMatRead_t_for_final_rare_common_gibbs_using_objects(
                eMatRead_t, P.rare_per_hap_info, P.common_snp_index, P.snp_is_common, sampleReads, P.hapMatcherR, grid, P.distinctHapsIE,
                P.special_helper, P.special_matrix, a->panel->ref_error, which_haps_to_use, rescale, a->Jmax, a->maxDifferenceBetweenReads,
                P.rare_per_snp_info);
        } else {
            Rcpp_make_eMatRead_t_for_gibbs_using_objects(eMatRead_t, sampleReads, P.hapMatcher, P.hapMatcherR, true, grid, P.rhb_t,
                                                         P.distinctHapsIE, P.special_helper, P.special_matrix, a->panel->ref_error,
                                                         which_haps_to_use, rescale, a->Jmax, a->maxDifferenceBetweenReads, true);
        }
        if (read_category) {
            arma::ivec number_of_non_1_reads(nReads), cat(nReads);
            arma::imat indices_of_non_1_reads(K, nReads);
            rcpp_evaluate_read_variability(eMatRead_t, number_of_non_1_reads, indices_of_non_1_reads, cat, 20);
            // the driver's overrides, gibbs-nipt.cpp (force_reset_read_category_zero / disable_read_category_usage)
            if (a->flags & QUILT_F_FORCE_RESET_READ_CATEGORY_0)
                for (int r = 0; r < nReads; ++r) if (cat(r) != 1) cat(r) = 0;
            if (a->flags & QUILT_F_DISABLE_READ_CATEGORY_USAGE) cat.fill(0);
            for (int r = 0; r < nReads; ++r) read_category[r] = cat(r);
        }
        return QUILT_OK;
    });
}

// panel words of the selected haplotypes.  The reference has no function returning packed words; what it has is
// the lookup (distinctHapsB row, or the special-matrix binary search, gibbs-small.cpp:204-231 / :579-597) and the
// expansion rcpp_int_expand.  This entry performs the lookup with the reference's own
// rcpp_simple_binary_matrix_search, expands every word with rcpp_int_expand and re-packs the bits, so both
// primitives are exercised; the all-SNP axis is assembled as rare_common.R:202-322 describes.
int quilt_ref_unpack_panel(const QuiltPanel* p, int32_t K, const int32_t* which, int32_t all_snps, uint32_t* words) {
    if (!p || !which || !words) return QUILT_ERR_BAD_ARG;
    return guarded([&]() -> int {
        Rcpp::IntegerMatrix helper = foreign_matrix<Rcpp::INTSXP>(p->eMatDH_special_matrix_helper, p->nGrids, 2);
        Rcpp::IntegerMatrix special = foreign_matrix<Rcpp::INTSXP>(p->eMatDH_special_matrix, p->n_special, 2);
        auto word_of = [&](int k0, int g) -> uint32_t {
            const int sym = p->hapMatcherR[(size_t)g * p->K_full + k0];
            int w;
            if (sym > 0) w = p->distinctHapsB[(size_t)g * p->nMaxDH + (sym - 1)];
            else w = rcpp_simple_binary_matrix_search(k0, special, helper(g, 0), helper(g, 1));
            arma::ivec hapc(1);
            hapc(0) = w;
            Rcpp::IntegerVector bits = rcpp_int_expand(hapc, 32);
            uint32_t out = 0;
            for (int b = 0; b < 32; ++b) if (bits[b]) out |= (1u << b);
            return out;
        };
        if (!all_snps) {
            for (int g = 0; g < p->nGrids; ++g)
                for (int k = 0; k < K; ++k) words[(size_t)g * K + k] = word_of(which[k] - 1, g);
            return QUILT_OK;
        }
        const int nG = (p->nSNPs_all + 31) / 32;
        std::fill(words, words + (size_t)nG * K, 0u);
        std::vector<uint32_t> cw((size_t)p->nGrids * K);
        for (int g = 0; g < p->nGrids; ++g)
            for (int k = 0; k < K; ++k) cw[(size_t)g * K + k] = word_of(which[k] - 1, g);
        for (int s = 0; s < p->nSNPs_all; ++s) {
            if (!p->snp_is_common[s]) continue;
            const int cs = p->common_snp_index[s] - 1;
            for (int k = 0; k < K; ++k)
                if ((cw[(size_t)(cs / 32) * K + k] >> (cs % 32)) & 1u) words[(size_t)(s / 32) * K + k] |= (1u << (s % 32));
        }
        for (int k = 0; k < K; ++k) {
            const int h = which[k] - 1;
            for (int64_t j = p->rare_hap_offsets[h]; j < p->rare_hap_offsets[h + 1]; ++j) {
                const int s = p->rare_hap_snps[j] - 1;
                words[(size_t)(s / 32) * K + k] |= (1u << (s % 32));
            }
        }
        return QUILT_OK;
    });
}

int quilt_ref_forward_backward(int32_t K, int32_t nGrids, const double* eMatGrid_in, const double* transMatRate_in, double* alphaHat_out,
                               double* betaHat_out, double* c_out) {
    return guarded([&]() -> int {
        arma::cube alphaMatCurrent_tc(K, nGrids - 1, 1);
        alphaMatCurrent_tc.fill(1 / double(K));
        arma::cube transMatRate_tc_H(const_cast<double*>(transMatRate_in), 2, nGrids - 1, 1, true);
        arma::mat priorCurrent_m(K, 1);
        priorCurrent_m.fill(1 / double(K));
        arma::mat eMatGrid_t(const_cast<double*>(eMatGrid_in), (arma::uword)K, (arma::uword)nGrids, true);
        arma::mat alphaHat_t(alphaHat_out, (arma::uword)K, (arma::uword)nGrids, false, true);
        arma::mat betaHat_t(betaHat_out, (arma::uword)K, (arma::uword)nGrids, false, true);
        alphaHat_t.zeros();
        betaHat_t.zeros();
        arma::rowvec c(nGrids);
        rcpp_initialize_gibbs_forward_backward(alphaMatCurrent_tc, transMatRate_tc_H, priorCurrent_m, 0, alphaHat_t, betaHat_t, c, eMatGrid_t,
                                               false, Rcpp::NumericVector(0), Rcpp::NumericVector(0));
        std::memcpy(c_out, c.memptr(), sizeof(double) * (size_t)nGrids);
        return QUILT_OK;
    });
}


// Rcpp_haploid_dosage_versus_refs with the production settings of functions.R:2034-2070 (use_eigen = TRUE -> the "version 3"
// forward / backward, always_normalize = FALSE, normalize_emissions = TRUE, hapMatcherR + special symbols)
int quilt_ref_haploid_dosage_versus_refs(const QuiltHaploidArgs* a, QuiltHaploidOut* o) {
    if (!a || !o || !a->panel || !a->gl || !a->transMatRate_t) return QUILT_ERR_BAD_ARG;
    return guarded([&]() -> int {
        const QuiltPanel* p = a->panel;
        const int K = p->K_full, T = p->nGrids, nS = p->nSNPs;
        PanelObjects P(p, false, 0, nullptr);
        arma::imat hapMatcher((arma::uword)K, 1);   // only its row count is read when hapMatcherR is in use (:2256)
        arma::mat gl(const_cast<double*>(a->gl), 2, (arma::uword)nS, true);
        arma::mat transMatRate_t(const_cast<double*>(a->transMatRate_t), 2, (arma::uword)(T - 1), true);
        arma::mat arma_alphaHat_t(1, 1);
        std::vector<double> alpha((size_t)K * T, 0.0);
        Eigen::Map<Eigen::MatrixXd> eigen_alphaHat_t(alpha.data(), K, T);
        const bool want_beta = (a->flags & QUILT_HF_RETURN_BETAHAT) != 0, want_gamma = (a->flags & QUILT_HF_RETURN_GAMMA) != 0;
        const bool want_dosage = (a->flags & QUILT_HF_RETURN_DOSAGE) != 0, want_best = (a->flags & QUILT_HF_GET_BEST_HAPS) != 0;
        arma::mat betaHat_t = want_beta ? arma::mat((arma::uword)K, (arma::uword)T) : arma::mat(1, 1);
        arma::mat gamma_t = want_gamma ? arma::mat((arma::uword)K, (arma::uword)T) : arma::mat(1, 1);
        arma::mat gammaSmall_t(1, 1);
        arma::rowvec c((arma::uword)T);
        c.fill(1.0);   // functions.R:2024
        Rcpp::NumericVector dosage(nS);
        Rcpp::IntegerVector cols(T);
        for (int g = 0; g < T; ++g) cols[g] = a->gammaSmall_cols_to_get ? a->gammaSmall_cols_to_get[g] : -1;
        Rcpp::List best(a->n_thinned > 0 ? a->n_thinned : 0);
        // eMatDH_special_grid_which: 1-based index of the grid among the grids that hold special haplotypes, else 0
        Rcpp::IntegerVector grid_which(T);
        int n_sp_grids = 0;
        for (int g = 0; g < T; ++g) {
            const int s1 = p->eMatDH_special_matrix_helper[g], e1 = p->eMatDH_special_matrix_helper[T + g];
            grid_which[g] = (s1 > 0 && e1 >= s1) ? ++n_sp_grids : 0;
        }
        Rcpp::List special_values_list(0);
        arma::imat rhb_t(1, 1);
        Rcpp_haploid_dosage_versus_refs(gl, arma_alphaHat_t, eigen_alphaHat_t, betaHat_t, c, gamma_t, gammaSmall_t, best, dosage, transMatRate_t, rhb_t,
                                        p->ref_error, true, P.distinctHapsB, P.distinctHapsIE, P.special_helper, P.special_matrix, true, hapMatcher,
                                        P.hapMatcherR, true, cols, grid_which, special_values_list, a->K_top_matches, 1,
                                        a->min_emission_prob_normalization_threshold, want_beta, want_dosage, want_gamma, false, want_best,
                                        false /*is_version_2*/, false, false, false /*always_normalize*/, true /*use_eigen*/, true /*normalize_emissions*/);
        if (o->dosage) std::memcpy(o->dosage, dosage.begin(), sizeof(double) * (size_t)nS);
        if (o->c) std::memcpy(o->c, c.memptr(), sizeof(double) * (size_t)T);
        if (o->alphaHat_t) std::memcpy(o->alphaHat_t, alpha.data(), sizeof(double) * alpha.size());
        if (o->betaHat_t && want_beta) std::memcpy(o->betaHat_t, betaHat_t.memptr(), sizeof(double) * (size_t)K * T);
        if (o->gamma_t && want_gamma) std::memcpy(o->gamma_t, gamma_t.memptr(), sizeof(double) * (size_t)K * T);
        if (want_best && o->best_haps_count) {
            for (int t = 0; t < a->n_thinned; ++t) {
                Rcpp::List e = Rcpp::as<Rcpp::List>(best[t]);
                Rcpp::IntegerVector tm = Rcpp::as<Rcpp::IntegerVector>(e["top_matches"]);
                Rcpp::NumericVector tv = Rcpp::as<Rcpp::NumericVector>(e["top_matches_values"]);
                o->best_haps_count[t] = tm.size();
                for (int i = 0; i < tm.size() && i < a->best_cap; ++i) {
                    if (o->best_haps) o->best_haps[(size_t)t * a->best_cap + i] = tm[i];
                    if (o->best_haps_values) o->best_haps_values[(size_t)t * a->best_cap + i] = tv[i];
                }
            }
        }
        return QUILT_OK;
    });
}

}  // extern "C"
