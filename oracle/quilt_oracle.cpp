// quilt_oracle.cpp — CPU ORACLE (test infrastructure, NOT product code).
//
// A statement-order restatement, in dependency-free C++17, of the reference's
// per-sample Gibbs hot path (rwdavies/QUILT @3c3a718, v2.0.4).  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this library; the product path (quilt_b200/csrc) never does.
//
// PARITY STATUS: pinned to the reference ITSELF.  oracle/_ref/libquiltref.so is the
// unmodified /root/reference/QUILT/src/{copied-from-stitch, gibbs-small, gibbs-nipt,
// gibbs-nipt-block}.cpp compiled against the header-only RcppArmadillo stand-in of
// oracle/refshim/; tests/test_ref_pins_oracle.py demands that this restatement and
// that library agree BIT FOR BIT (labels, H_class, read categories, alpha / beta /
// eMatGrid / c, hapProbs / genProbs, per-sweep likelihood table) on diploid and NIPT
// calls, and the golden fixtures in tests/golden/ are the reference's outputs.  What
// the stand-in cannot verify is real Armadillo's accumulation order (two interleaved
// accumulators is what arrayops::accumulate / accu_proxy_linear do in non-fast-math
// builds; stated below) — last-ulp differences there are inside the 1e-4 DS/GP
// tolerance of north_star and do not reach the label decisions.  The reference
// tests' own invariants are re-expressed in tests/test_oracle_invariants.py.
//
// Conventions that matter for bit-level agreement with the reference:
//  * Armadillo evaluates element-wise expression templates one element at a
//    time with each binary op rounded separately (the reference is built for
//    baseline x86-64: QUILT/src/Makevars has -O3 but no -march, so no FMA
//    contraction).  Build this file with -ffp-contract=off.
//  * sum()/accu() of a vector expression uses two interleaved accumulators
//    (even / odd elements), val1 + val2 at the end (Armadillo's
//    arrayops::accumulate / accu_proxy_linear).  Rcpp sugar sum() is a plain
//    left-to-right loop.
//  * mat/cube/vec constructors zero-fill (Armadillo >= 10.5).
//
// Each function cites the reference file:line it follows.

#include "quilt_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <numeric>
#include <vector>

namespace {

// ---------------------------------------------------------------- small containers
struct Mat {  // column-major, zero-filled like arma::mat(n_rows, n_cols)
    int nr = 0, nc = 0;
    std::vector<double> d;
    Mat() {}
    Mat(int r, int c, double v = 0.0) : nr(r), nc(c), d((size_t)r * c, v) {}
    double* col(int j) { return d.data() + (size_t)j * nr; }
    const double* col(int j) const { return d.data() + (size_t)j * nr; }
    double& operator()(int i, int j) { return d[(size_t)j * nr + i]; }
    double operator()(int i, int j) const { return d[(size_t)j * nr + i]; }
    void fill(double v) { std::fill(d.begin(), d.end(), v); }
};
typedef std::vector<double> Vec;
typedef std::vector<int> IVec;

// Armadillo accu(): two interleaved accumulators (arrayops::accumulate, accu_proxy_linear)
template <class F>
inline double accu2(int n, F f) {
    double v1 = 0.0, v2 = 0.0;
    int i, j;
    for (i = 0, j = 1; j < n; i += 2, j += 2) {
        v1 += f(i);
        v2 += f(j);
    }
    if (i < n) v1 += f(i);
    return v1 + v2;
}
inline double accu2(const double* x, int n) {
    return accu2(n, [&](int i) { return x[i]; });
}
inline bool is_finite_d(double x) { return std::isfinite(x); }

// Rcpp::max on a NumericVector: returns the first NA/NaN met (sugar/functions/max.h)
inline double rcpp_max(const double* x, int n) {
    if (n == 0) return -std::numeric_limits<double>::infinity();
    double m = x[0];
    if (std::isnan(m)) return m;
    for (int i = 1; i < n; i++) {
        double c = x[i];
        if (std::isnan(c)) return c;
        if (c > m) m = c;
    }
    return m;
}

struct Reads {
    int nReads;
    const int32_t *off, *u, *bq, *wif0;
};

// ---------------------------------------------------------------- gibbs-small.cpp:69-105
// rcpp_simple_binary_matrix_search; mat is [n x 2] column-major, s1/e1 1-based rows
int simple_binary_matrix_search(int val, const int32_t* mat, int nrow, int s1, int e1) {
    int nori = e1 - s1 + 1;
    if (nori == 1) return 0;  // sic: reference returns 0, not the word (gibbs-small.cpp:76-78)
    int n = nori;
    int i = n / 2;
    n = n / 4;
    int c = 0;
    while (c < 100) {
        c++;
        int v = mat[(size_t)0 * nrow + (s1 - 1 + i)];
        if (v == val) {
            return mat[(size_t)1 * nrow + (s1 - 1 + i)];
        } else if (v < val) {
            i += n;
        } else {
            i -= n;
        }
        n = n / 2;
        if (n < 1) n = 1;
        if (i < 0) i = 0;
        if (i > (nori - 1)) i = nori - 1;
    }
    return mat[(size_t)1 * nrow + s1];
}

// the 32-bit word of haplotype k0 (0-based, panel index) at (common) grid g:
// distinctHapsB lookup or special-matrix search (gibbs-small.cpp:579-597)
inline int32_t panel_word(const QuiltPanel* p, int k0, int g) {
    int kk = p->hapMatcherR[(size_t)g * p->K_full + k0];
    if (kk > 0) return p->distinctHapsB[(size_t)g * p->nMaxDH + (kk - 1)];
    return simple_binary_matrix_search(k0, p->eMatDH_special_matrix, p->n_special,
                                       p->eMatDH_special_matrix_helper[g],
                                       p->eMatDH_special_matrix_helper[(size_t)p->nGrids + g]);
}

// rare_per_snp_info for this call (rare_common.R:313-322): for each all-SNP index,
// the 1-based positions within which_haps_to_use that carry the alt allele
// (reference stores c(-1, k...) — we store only the k's; "length()==1" <=> empty here)
void build_rare_per_snp(const QuiltPanel* p, int K, const int32_t* which, std::vector<std::vector<int>>& out) {
    out.assign(p->nSNPs_all, std::vector<int>());
    for (int k = 1; k <= K; k++) {
        int h = which[k - 1] - 1;
        for (int64_t j = p->rare_hap_offsets[h]; j < p->rare_hap_offsets[h + 1]; j++) {
            int snp = p->rare_hap_snps[j];  // 1-based
            out[snp - 1].push_back(k);
        }
    }
}

// ---------------------------------------------------------------- gibbs-small.cpp:235-262 / copied-from-stitch.cpp:193-226
void rescale_eMatRead_col(double* e, int K, double d2) {
    double x = 0;
    for (int k = 0; k < K; k++)
        if (e[k] > x) x = e[k];
    double d1 = 1 / x;
    const double inf = std::numeric_limits<double>::infinity();
    // (x == R_NaN) is never true; kept for the shape of the original test
    if ((x == inf) | (x == -inf) | (x == 0) | (d1 == inf) | (d1 == -inf)) {
        for (int k = 0; k < K; k++) e[k] = 1;
    } else {
        for (int k = 0; k < K; k++) {
            e[k] *= d1;
            if (e[k] < d2) e[k] = d2;
        }
    }
}

// ---------------------------------------------------------------- gibbs-small.cpp:116-265
// Rcpp_make_eMatRead_t_for_gibbs_using_objects with use_hapMatcherR = TRUE,
// use_eMatDH_special_symbols = TRUE (QUILT2 production, quilt-prepare-reference.R:430-446)
void make_eMatRead_t_using_objects(Mat& eMatRead_t, const Reads& R, const QuiltPanel* p, const int32_t* which,
                                   bool rescale, int Jmax, double maxDifferenceBetweenReads) {
    const int nReads = R.nReads;
    const int K = eMatRead_t.nr;
    std::vector<int> haps_at_gridR(K);
    double eps, pR = 1, pA = 1, e;
    double d2 = 1 / maxDifferenceBetweenReads;
    int s1 = 0, e1 = 0;
    const double ref_error = p->ref_error;
    for (int iRead = 0; iRead < nReads; iRead++) {
        int J = R.off[iRead + 1] - R.off[iRead] - 1;
        const int32_t* bq = R.bq + R.off[iRead];
        const int32_t* u = R.u + R.off[iRead];
        int iGrid0 = u[0] / 32;  // grid(u(0)); production grid is floor(i/32) (quilt-prepare-reference.R:376-380)
        for (int k = 0; k < K; k++) haps_at_gridR[k] = p->hapMatcherR[(size_t)iGrid0 * p->K_full + (which[k] - 1)];
        s1 = p->eMatDH_special_matrix_helper[iGrid0];
        e1 = p->eMatDH_special_matrix_helper[(size_t)p->nGrids + iGrid0];
        int iGrid0_prev = iGrid0;
        if (J >= Jmax) J = Jmax;
        double* col = eMatRead_t.col(iRead);
        for (int j = 0; j <= J; j++) {
            if (bq[j] < 0) {
                eps = std::pow(10, (double(bq[j]) / 10));
                pR = 1 - eps;
                pA = eps / 3;
            }
            if (bq[j] > 0) {
                eps = std::pow(10, (-double(bq[j]) / 10));
                pR = eps / 3;
                pA = 1 - eps;
            }
            iGrid0 = u[j] / 32;
            if (iGrid0 != iGrid0_prev) {
                for (int k = 0; k < K; k++)
                    haps_at_gridR[k] = p->hapMatcherR[(size_t)iGrid0 * p->K_full + (which[k] - 1)];
                s1 = p->eMatDH_special_matrix_helper[iGrid0];
                e1 = p->eMatDH_special_matrix_helper[(size_t)p->nGrids + iGrid0];
            }
            iGrid0_prev = iGrid0;
            for (int k = 0; k < K; k++) {
                if (haps_at_gridR[k] > 0) {
                    e = p->distinctHapsIE[(size_t)u[j] * p->nMaxDH + (haps_at_gridR[k] - 1)];
                } else {
                    int bvtd = simple_binary_matrix_search(which[k] - 1, p->eMatDH_special_matrix, p->n_special, s1, e1);
                    uint32_t tmp = (uint32_t)bvtd;
                    if (((tmp >> (u[j] % 32)) & 0x1) == 1) {
                        e = 1 - ref_error;
                    } else {
                        e = ref_error;
                    }
                }
                col[k] *= (e * pA + (1 - e) * pR);
            }
        }
        if (rescale) rescale_eMatRead_col(col, K, d2);
    }
}

// ---------------------------------------------------------------- gibbs-small.cpp:275-465
// Rcpp_make_eMatRead_t_for_final_rare_common_gibbs_using_objects
void make_eMatRead_t_rare_common(Mat& eMatRead_t, const Reads& R, const QuiltPanel* p, const int32_t* which,
                                 bool rescale, int Jmax, double maxDifferenceBetweenReads,
                                 const std::vector<std::vector<int>>& rare_per_snp) {
    const int nReads = R.nReads;
    const int K = eMatRead_t.nr;
    std::vector<int> haps_at_gridR(K);
    double eps, pR = 1, pA = 1, e;
    double d2 = 1 / maxDifferenceBetweenReads;
    int s1 = 0, e1 = 0;
    const double ref_error = p->ref_error;
    const double one_minus_ref_error = 1 - ref_error;
    double xe1 = 1, xe2 = 1;
    for (int iRead = 0; iRead < nReads; iRead++) {
        int J = R.off[iRead + 1] - R.off[iRead] - 1;
        const int32_t* bq = R.bq + R.off[iRead];
        const int32_t* u = R.u + R.off[iRead];
        int iGrid0_prev = -1;
        if (J >= Jmax) J = Jmax;
        double* col = eMatRead_t.col(iRead);
        for (int j = 0; j <= J; j++) {
            if (bq[j] < 0) {
                eps = std::pow(10, (double(bq[j]) / 10));
                pR = 1 - eps;
                pA = eps / 3;
            }
            if (bq[j] > 0) {
                eps = std::pow(10, (-double(bq[j]) / 10));
                pR = eps / 3;
                pA = 1 - eps;
            }
            if (p->snp_is_common[u[j]]) {
                int u_j_common = p->common_snp_index[u[j]] - 1;
                int iGrid0 = u_j_common / 32;
                if (iGrid0 != iGrid0_prev) {
                    for (int k = 0; k < K; k++)
                        haps_at_gridR[k] = p->hapMatcherR[(size_t)iGrid0 * p->K_full + (which[k] - 1)];
                    s1 = p->eMatDH_special_matrix_helper[iGrid0];
                    e1 = p->eMatDH_special_matrix_helper[(size_t)p->nGrids + iGrid0];
                }
                iGrid0_prev = iGrid0;
                for (int k = 0; k < K; k++) {
                    if (haps_at_gridR[k] > 0) {
                        e = p->distinctHapsIE[(size_t)u_j_common * p->nMaxDH + (haps_at_gridR[k] - 1)];
                    } else {
                        int bvtd = simple_binary_matrix_search(which[k] - 1, p->eMatDH_special_matrix, p->n_special, s1, e1);
                        uint32_t tmp = (uint32_t)bvtd;
                        if (((tmp >> (u_j_common % 32)) & 0x1) == 1) {
                            e = 1 - ref_error;
                        } else {
                            e = ref_error;
                        }
                    }
                    col[k] *= (e * pA + (1 - e) * pR);
                }
            } else {
                const std::vector<int>& k_with_alt = rare_per_snp[u[j]];
                if (k_with_alt.empty()) {
                    if (!rescale) {
                        xe1 = ref_error * pA + one_minus_ref_error * pR;
                        for (int k = 0; k < K; k++) col[k] *= xe1;
                    }
                } else {
                    xe1 = ref_error * pA + one_minus_ref_error * pR;
                    xe2 = one_minus_ref_error * pA + ref_error * pR;
                    for (int k = 0; k < K; k++) col[k] *= xe1;
                    for (size_t ik = 0; ik < k_with_alt.size(); ik++) {
                        int k = k_with_alt[ik] - 1;
                        col[k] *= xe2 / xe1;
                    }
                }
            }
        }
        if (rescale) rescale_eMatRead_col(col, K, d2);
    }
}

// ---------------------------------------------------------------- gibbs-nipt.cpp:338-382
void evaluate_read_variability(const Mat& eMatRead_t, IVec& number_of_non_1_reads, std::vector<int32_t>& indices_of_non_1_reads,
                               IVec& read_category) {
    int K = eMatRead_t.nr;
    int nReads = eMatRead_t.nc;
    double thresh = 1 - std::pow(10, -12);
    int thresh2 = K * 0.20;
    std::fill(read_category.begin(), read_category.end(), 0);
    for (int iRead = 0; iRead < nReads; iRead++) {
        int c = 0;
        double val = -1;
        bool more_than_two = false;
        const double* e = eMatRead_t.col(iRead);
        int32_t* idx = indices_of_non_1_reads.data() + (size_t)iRead * K;
        for (int k = 0; k < K; k++) {
            if (e[k] < thresh) {
                idx[c] = k;
                c++;
                if (val == -1) {
                    val = e[k];
                } else {
                    if (val != e[k]) more_than_two = true;
                }
            }
        }
        number_of_non_1_reads[iRead] = c - 1 + 1;
        if (number_of_non_1_reads[iRead] == 0) {
            read_category[iRead] = 1;
        } else if (!more_than_two) {
            read_category[iRead] = 2;
        } else if (number_of_non_1_reads[iRead] < thresh2) {
            read_category[iRead] = 3;
        } else {
            read_category[iRead] = 0;
        }
    }
}

// ---------------------------------------------------------------- copied-from-stitch.cpp:234-310 (bound = false)
void make_eMatGrid_t(Mat& eMatGrid_t, const Mat& eMatRead_t, const IVec& H, const Reads& R, int hap) {
    const int K = eMatRead_t.nr;
    for (int iRead = 0; iRead < R.nReads; iRead++) {
        if (H[iRead] == hap) {
            int w = R.wif0[iRead];
            double* g = eMatGrid_t.col(w);
            const double* e = eMatRead_t.col(iRead);
            for (int k = 0; k < K; k++) g[k] *= e[k];
        }
    }
}

// transition accessors: transMatRate_tc_H(row, g, s=0)
struct Trans {
    const double* t;
    double operator()(int row, int g) const { return t[(size_t)g * 2 + row]; }
};

// ---------------------------------------------------------------- copied-from-stitch.cpp:340-387
// Rcpp_run_forward_haploid; priorCurrent_m = alphaMatCurrent_tc = 1/K (quilt.R:756-757)
void run_forward_haploid(Mat& alphaHat_t, Vec& c, const Mat& eMatGrid_t, double one_over_K_prior, const Trans& tm,
                         bool initialize_only) {
    const int K = alphaHat_t.nr;
    const int nGrids = alphaHat_t.nc;
    double* a0 = alphaHat_t.col(0);
    const double* e0 = eMatGrid_t.col(0);
    for (int k = 0; k < K; k++) a0[k] = one_over_K_prior * e0[k];
    c[0] = 1 / accu2(a0, K);
    for (int k = 0; k < K; k++) a0[k] = a0[k] * c[0];
    if (initialize_only) return;
    for (int iGrid = 1; iGrid < nGrids; iGrid++) {
        double* a = alphaHat_t.col(iGrid);
        const double* ap = alphaHat_t.col(iGrid - 1);
        const double* e = eMatGrid_t.col(iGrid);
        const double t0 = tm(0, iGrid - 1), t1 = tm(1, iGrid - 1);
        for (int k = 0; k < K; k++) a[k] = e[k] * (t0 * ap[k] + t1 * one_over_K_prior);
        c[iGrid] = 1 / accu2(a, K);
        for (int k = 0; k < K; k++) a[k] *= c[iGrid];
    }
}

// ---------------------------------------------------------------- copied-from-stitch.cpp:392-409
void run_backward_haploid(Mat& betaHat_t, const Vec& c, const Mat& eMatGrid_t, double alphaMat_const, const Trans& tm) {
    const int K = betaHat_t.nr;
    const int nGrids = eMatGrid_t.nc;
    Vec e_times_b(K);
    for (int iGrid = nGrids - 2; iGrid >= 0; --iGrid) {
        const double* e = eMatGrid_t.col(iGrid + 1);
        const double* b1 = betaHat_t.col(iGrid + 1);
        for (int k = 0; k < K; k++) e_times_b[k] = e[k] * b1[k];
        double x = tm(1, iGrid) * accu2(K, [&](int k) { return alphaMat_const * e_times_b[k]; });
        double* b = betaHat_t.col(iGrid);
        const double t0 = tm(0, iGrid);
        for (int k = 0; k < K; k++) b[k] = c[iGrid] * (x + t0 * e_times_b[k]);
    }
}

// ---------------------------------------------------------------- copied-from-stitch.cpp:417-440
void run_backward_haploid_QUILT_faster(Mat& betaHat_t, const Vec& c, const Mat& eMatGrid_t, const Trans& tm,
                                       const std::vector<uint8_t>& grid_has_read) {
    const int K = betaHat_t.nr;
    const int nGrids = eMatGrid_t.nc;
    const double one_over_K = 1 / double(K);
    Vec e_times_b(K);
    for (int iGrid = nGrids - 2; iGrid >= 0; --iGrid) {
        double* b = betaHat_t.col(iGrid);
        const double* b1 = betaHat_t.col(iGrid + 1);
        const double t0 = tm(0, iGrid);
        if (grid_has_read[iGrid + 1]) {
            const double* e = eMatGrid_t.col(iGrid + 1);
            for (int k = 0; k < K; k++) e_times_b[k] = e[k] * b1[k];
            double x = tm(1, iGrid) * accu2(e_times_b.data(), K) * one_over_K;
            for (int k = 0; k < K; k++) b[k] = c[iGrid] * (x + t0 * e_times_b[k]);
        } else {
            double x = tm(1, iGrid) * accu2(b1, K) * one_over_K;
            for (int k = 0; k < K; k++) b[k] = c[iGrid] * (x + t0 * b1[k]);
        }
    }
}

// ---------------------------------------------------------------- gibbs-nipt.cpp:630-661 (alphaMatCurrent_tc = 1/K)
void alpha_forward_one(int iGrid, int K, Mat& alphaHat_t, const Trans& tm, const Mat& eMatGrid_t, double alphaMat_const,
                       Vec& c, double& minus_log_c_sum, bool normalize) {
    int g1 = iGrid - 1;
    const double* ap = alphaHat_t.col(g1);
    double alphaConst = tm(1, g1) * accu2(ap, K);
    double x = tm(0, g1);
    double c2 = c[iGrid];
    double* a = alphaHat_t.col(iGrid);
    const double* e = eMatGrid_t.col(iGrid);
    for (int k = 0; k < K; k++) a[k] = c2 * e[k] * (x * ap[k] + alphaConst * alphaMat_const);
    if (normalize) {
        double s = 1 / accu2(a, K);
        minus_log_c_sum -= std::log(s);
        c[iGrid] *= s;
        for (int k = 0; k < K; k++) a[k] *= s;
    }
}

// ---------------------------------------------------------------- gibbs-nipt.cpp:671-707
void alpha_forward_one_QUILT_faster(int iGrid, int K, Mat& alphaHat_t, const Trans& tm, const Mat& eMatGrid_t, Vec& c,
                                    double& minus_log_c_sum, const std::vector<uint8_t>& grid_has_read, bool normalize) {
    int g1 = iGrid - 1;
    const double one_over_K = 1 / double(K);
    const double* ap = alphaHat_t.col(g1);
    double alphaConst = tm(1, g1) * accu2(ap, K);
    double x = tm(0, g1);
    double c2 = c[iGrid];
    double* a = alphaHat_t.col(iGrid);
    const double jump = alphaConst * one_over_K;
    if (grid_has_read[iGrid]) {
        const double* e = eMatGrid_t.col(iGrid);
        for (int k = 0; k < K; k++) a[k] = e[k] * (x * ap[k] + jump);
    } else {
        for (int k = 0; k < K; k++) a[k] = (x * ap[k] + jump);
    }
    if (normalize) {
        double s = 1 / (c2 * accu2(a, K));
        minus_log_c_sum -= std::log(s);
        c[iGrid] *= s;
        s *= c2;
        for (int k = 0; k < K; k++) a[k] *= s;
    }
}

// ---------------------------------------------------------------- gibbs-nipt.cpp:712-727
void reinitialize_in_iterations(Mat& alphaHat_t, Vec& c, double prior, const Mat& eMatGrid_t, int K) {
    double* a = alphaHat_t.col(0);
    const double* e = eMatGrid_t.col(0);
    for (int k = 0; k < K; k++) a[k] = prior * e[k];
    c[0] = 1 / accu2(a, K);
    for (int k = 0; k < K; k++) a[k] *= c[0];
}

// ---------------------------------------------------------------- gibbs-nipt-block.cpp:146-164
double get_log_p_H_class(const IVec& H_class, double ff) {
    double out = 0;
    double vals[8];
    vals[0] = 0;
    vals[1] = std::log(0.5);
    vals[2] = std::log(0.5 - ff * 0.5);
    vals[3] = std::log(ff * 0.5);
    vals[4] = std::log(1.0 - ff * 0.5);
    vals[5] = std::log(0.5 + ff * 0.5);
    vals[6] = std::log(0.5);
    vals[7] = 0;
    for (size_t i = 0; i < H_class.size(); i++) out += vals[H_class[i]];
    return out;
}

// ---------------------------------------------------------------- gibbs-nipt-block.cpp:169-208
double get_log_p_H_class2(int n1, int n2, int n3, int n4, int n5, int n6, double ff) {
    double out = 0;
    if (ff == 0) {
        out = 0 + n1 * std::log(0.5) + n2 * std::log(0.5 - ff * 0.5) + n3 * std::log(0.001) + n4 * std::log(1 - ff * 0.5) +
              n5 * std::log(1 * 0.5 + ff * 0.5) + n6 * std::log(1 * 0.5);
    } else if (ff == 1) {
        out = 0 + n1 * std::log(0.5) + n2 * std::log(0.001) + n3 * std::log(ff * 0.5) + n4 * std::log(1 - ff * 0.5) +
              n5 * std::log(1 * 0.5 + ff * 0.5) + n6 * std::log(1 * 0.5);
    } else {
        out = 0 + n1 * std::log(0.5) + n2 * std::log(0.5 - ff * 0.5) + n3 * std::log(ff * 0.5) + n4 * std::log(1 - ff * 0.5) +
              n5 * std::log(1 * 0.5 + ff * 0.5) + n6 * std::log(1 * 0.5);
    }
    return out;
}

// ---------------------------------------------------------------- gibbs-nipt.cpp:1463-1477
double calc_prob_of_set_of_reads(double ff, const double rc[3]) {
    const double prior_probs[3] = {0.5, (1 - ff) / 2, (ff / 2)};
    int n = rc[0] + rc[1] + rc[2];
    double r = std::lgamma(1.0 * (n + 1.0));
    for (int i = 0; i < 3; i++) {
        if (prior_probs[i] > 0) {
            r += rc[i] * std::log(prior_probs[i]) - std::lgamma(1.0 * (rc[i] + 1.0));
        }
    }
    return r;
}

// ---------------------------------------------------------------- gibbs-nipt.cpp:1483-1519
void calculate_likelihoods_values(const Vec& c1, const Vec& c2, const Vec& c3, const IVec& H, int nGrids,
                                  const double prior_probs[3], double ff, double to_out[7]) {
    double d1 = 0, d2 = 0, d3 = 0, dH = 0;
    for (int iGrid = 0; iGrid < nGrids; iGrid++) {
        d1 -= std::log(c1[iGrid]);
        d2 -= std::log(c2[iGrid]);
        d3 -= std::log(c3[iGrid]);
    }
    for (size_t iRead = 0; iRead < H.size(); iRead++) dH += std::log(prior_probs[H[iRead] - 1]);
    to_out[0] = d1;
    to_out[1] = d2;
    to_out[2] = d3;
    to_out[3] = d1 + d2 + d3;
    to_out[4] = dH;
    to_out[5] = to_out[3] + to_out[4];
    double rc[3] = {0, 0, 0};
    for (size_t iRead = 0; iRead < H.size(); iRead++) rc[H[iRead] - 1] += 1;
    to_out[6] = calc_prob_of_set_of_reads(ff, rc);
}

// full per-call state (the reference passes these as a dozen arma::mat& arguments)
struct State {
    int K, nGrids, nReads, nHaps;
    bool sample_is_diploid;
    double ff;
    Mat alphaHat_t[3], betaHat_t[3], eMatGrid_t[3];
    Vec c[3];
    Mat eMatRead_t;
    IVec H, H_class;
    IVec number_of_non_1_reads, read_category;
    std::vector<int32_t> indices_of_non_1_reads;
    std::vector<uint8_t> grid_has_read;
    Trans tm;
    Reads R;
    double prior_probs[3];
    double rlc[7][3];
};

// ---------------------------------------------------------------- gibbs-nipt.cpp:733-1295
void sample_reads_in_grid(State& S, int& iRead, int iGrid, bool& done_reads, int& read_wif_iRead, double pC[3], double pA1[3],
                          double pA2[3], Mat& alphaHat_m, Mat& betaHat_m, Mat& ab_m, const double* runif_reads, int iteration,
                          bool record_read_set, double class_sum_cutoff, bool gibbs_initialize_iteratively,
                          int first_read_for_gibbs_initialization) {
    const int K = S.K;
    const int nReads = S.nReads;
    const bool sample_is_diploid = S.sample_is_diploid;
    int h_rC = 0, h_rA1 = 1, h_rA2 = 2, h_rN = 0;
    int k = 0, ik;
    double prod_pC, prod_pA1, prod_pA2, norm_pC, norm_pA1, norm_pA2;
    double val1, val2, val3;
    double cumsum_flip_probs[3];
    double denom, chance, alphaConst;
    const double* e = nullptr;  // eMatRead_t_col
    bool this_grid_has_at_least_one_read = false;
    bool at_least_one_read_has_changed = false;
    bool currently_doing_normal_progress = false;
    bool currently_doing_gibbs_initialization = false;
    bool currently_doing_pass_through = false;
    while ((done_reads == false) & ((read_wif_iRead) == iGrid)) {
        if (!sample_is_diploid || (sample_is_diploid && S.read_category[iRead] != 1)) {
            if (!gibbs_initialize_iteratively) {
                currently_doing_normal_progress = true;
            } else {
                if ((iRead < first_read_for_gibbs_initialization) & (iteration == 0)) {
                    currently_doing_pass_through = true;
                } else if ((first_read_for_gibbs_initialization <= iRead) & (iteration == 0)) {
                    currently_doing_pass_through = false;
                    currently_doing_gibbs_initialization = true;
                } else if ((iRead < first_read_for_gibbs_initialization) & (iteration == 1)) {
                    currently_doing_pass_through = false;
                    currently_doing_gibbs_initialization = true;
                } else {
                    currently_doing_gibbs_initialization = false;
                    currently_doing_normal_progress = true;
                }
            }
            if (!this_grid_has_at_least_one_read) {
                for (int j = 0; j < 3; j++) pC[j] = pA1[j] = pA2[j] = 1;
                for (int h = 0; h < S.nHaps; h++) {
                    const double* a = S.alphaHat_t[h].col(iGrid);
                    const double* b = S.betaHat_t[h].col(iGrid);
                    double* am = alphaHat_m.col(h);
                    double* bm = betaHat_m.col(h);
                    double* ab = ab_m.col(h);
                    for (int kk = 0; kk < K; kk++) {
                        am[kk] = a[kk];
                        bm[kk] = b[kk];
                        ab[kk] = am[kk] * bm[kk];
                    }
                }
                pC[0] = accu2(ab_m.col(0), K);
                pC[1] = accu2(ab_m.col(1), K);
                if (!sample_is_diploid) pC[2] = accu2(ab_m.col(2), K);
                this_grid_has_at_least_one_read = true;
            }
            if (currently_doing_normal_progress) {
                e = S.eMatRead_t.col(iRead);
                h_rC = S.H[iRead] - 1;
                if (h_rC == 0) {
                    h_rA1 = 1;
                    h_rA2 = 2;
                } else if (h_rC == 1) {
                    h_rA1 = 0;
                    h_rA2 = 2;
                } else if (h_rC == 2) {
                    h_rA1 = 0;
                    h_rA2 = 1;
                }
                for (int j = 0; j < 3; j++) {
                    pA1[j] = pC[j];
                    pA2[j] = pC[j];
                }
                const int cat = S.read_category[iRead];
                const int nn1 = S.number_of_non_1_reads[iRead];
                const int32_t* idx = S.indices_of_non_1_reads.data() + (size_t)iRead * K;
                if (cat == 0) {
                    const double* abC = ab_m.col(h_rC);
                    const double* abA1 = ab_m.col(h_rA1);
                    pA1[h_rC] = accu2(K, [&](int i) { return abC[i] / e[i]; });
                    pA1[h_rA1] = accu2(K, [&](int i) { return abA1[i] * e[i]; });
                    if (!sample_is_diploid) {
                        const double* abA2 = ab_m.col(h_rA2);
                        pA2[h_rA2] = accu2(K, [&](int i) { return abA2[i] * e[i]; });
                    }
                } else if (cat == 2) {
                    val1 = val2 = val3 = 0;
                    if (sample_is_diploid) {
                        for (ik = 0; ik < nn1; ik++) {
                            k = idx[ik];
                            val1 += ab_m(k, h_rC);
                            val2 += ab_m(k, h_rA1);
                        }
                        pA1[h_rC] += val1 * (1 / e[k] - 1);
                        pA1[h_rA1] += val2 * (e[k] - 1);
                    } else {
                        for (ik = 0; ik < nn1; ik++) {
                            k = idx[ik];
                            val1 += ab_m(k, h_rC);
                            val2 += ab_m(k, h_rA1);
                            val3 += ab_m(k, h_rA2);
                        }
                        pA1[h_rC] += val1 * (1 / e[k] - 1);
                        pA1[h_rA1] += val2 * (e[k] - 1);
                        pA2[h_rA2] += val3 * (e[k] - 1);
                    }
                } else if (cat == 3) {
                    if (sample_is_diploid) {
                        for (ik = 0; ik < nn1; ik++) {
                            k = idx[ik];
                            pA1[h_rC] += ab_m(k, h_rC) * (1 / e[k] - 1);
                            pA1[h_rA1] += ab_m(k, h_rA1) * (e[k] - 1);
                        }
                    } else {
                        for (ik = 0; ik < nn1; ik++) {
                            k = idx[ik];
                            pA1[h_rC] += ab_m(k, h_rC) * (1 / e[k] - 1);
                            pA1[h_rA1] += ab_m(k, h_rA1) * (e[k] - 1);
                            pA2[h_rA2] += ab_m(k, h_rA2) * (e[k] - 1);
                        }
                    }
                }
                pA2[h_rA1] = pC[h_rA1];
                pA2[h_rC] = pA1[h_rC];
            } else if (currently_doing_gibbs_initialization) {
                e = S.eMatRead_t.col(iRead);
                h_rC = 0;
                h_rA1 = 1;
                h_rA2 = 2;
                for (int j = 0; j < 3; j++) {
                    pA1[j] = pC[j];
                    pA2[j] = pC[j];
                }
                const double* ab0 = ab_m.col(h_rC);
                const double* ab1 = ab_m.col(h_rA1);
                pC[h_rC] = accu2(K, [&](int i) { return ab0[i] * e[i]; });
                pA1[h_rA1] = accu2(K, [&](int i) { return ab1[i] * e[i]; });
                if (!sample_is_diploid) {
                    const double* ab2 = ab_m.col(h_rA2);
                    pA2[h_rA2] = accu2(K, [&](int i) { return ab2[i] * e[i]; });
                }
            } else {
                for (int j = 0; j < 3; j++) {
                    pA1[j] = pC[j];
                    pA2[j] = pC[j];
                }
            }
            prod_pC = (pC[0] * pC[1] * pC[2]) * S.prior_probs[h_rC];
            prod_pA1 = (pA1[0] * pA1[1] * pA1[2]) * S.prior_probs[h_rA1];
            prod_pA2 = (pA2[0] * pA2[1] * pA2[2]) * S.prior_probs[h_rA2];
            denom = prod_pC + prod_pA1 + prod_pA2;
            norm_pC = prod_pC / denom;
            norm_pA1 = prod_pA1 / denom;
            norm_pA2 = prod_pA2 / denom;
            chance = runif_reads[(size_t)nReads * iteration + iRead];
            cumsum_flip_probs[0] = cumsum_flip_probs[1] = cumsum_flip_probs[2] = 0;
            cumsum_flip_probs[h_rC] = norm_pC;
            cumsum_flip_probs[h_rA1] = norm_pA1;
            cumsum_flip_probs[h_rA2] = norm_pA2;
            cumsum_flip_probs[1] += cumsum_flip_probs[0];
            cumsum_flip_probs[2] += cumsum_flip_probs[1];
            h_rN = 0;
            for (int i = 2; i >= 0; i--) {
                if (chance < cumsum_flip_probs[i]) h_rN = i;
            }
            if (((h_rN != h_rC) | currently_doing_gibbs_initialization) & (!currently_doing_pass_through)) {
                at_least_one_read_has_changed = true;
                S.H[iRead] = h_rN + 1;
                if (currently_doing_normal_progress) {
                    double* am = alphaHat_m.col(h_rC);
                    double* ab = ab_m.col(h_rC);
                    for (int kk = 0; kk < K; kk++) am[kk] /= e[kk];
                    for (int kk = 0; kk < K; kk++) ab[kk] /= e[kk];
                }
                {
                    double* am = alphaHat_m.col(h_rN);
                    double* ab = ab_m.col(h_rN);
                    for (int kk = 0; kk < K; kk++) am[kk] *= e[kk];
                    for (int kk = 0; kk < K; kk++) ab[kk] *= e[kk];
                }
                if (currently_doing_normal_progress) {
                    if (h_rC < 2 || !sample_is_diploid) {
                        double* g = S.eMatGrid_t[h_rC].col(iGrid);
                        for (int kk = 0; kk < K; kk++) g[kk] /= e[kk];
                    }
                }
                if (h_rN < 2 || !sample_is_diploid) {
                    double* g = S.eMatGrid_t[h_rN].col(iGrid);
                    for (int kk = 0; kk < K; kk++) g[kk] *= e[kk];
                }
                if (currently_doing_normal_progress) {
                    for (int i = 0; i < 3; i++) {
                        if (h_rC == 0) {
                            if (h_rN == 1) pC[i] = pA1[i];
                            if (h_rN == 2) pC[i] = pA2[i];
                        } else if (h_rC == 1) {
                            if (h_rN == 0) pC[i] = pA1[i];
                            if (h_rN == 2) pC[i] = pA2[i];
                        } else if (h_rC == 2) {
                            if (h_rN == 0) pC[i] = pA1[i];
                            if (h_rN == 1) pC[i] = pA2[i];
                        }
                    }
                } else if (currently_doing_gibbs_initialization) {
                    for (int i = 0; i < 3; i++) {
                        if (h_rN == 1) pC[i] = pA1[i];
                        if (h_rN == 2) pC[i] = pA2[i];
                    }
                }
            }
            if (record_read_set) {
                double x[3] = {0, 0, 0};
                x[h_rC] = norm_pC;
                x[h_rA1] = norm_pA1;
                x[h_rA2] = norm_pA2;
                double local_min = 2;
                int local_min_which = 8;
                for (int i = 0; i < 7; i++) {
                    double y = std::abs(S.rlc[i][0] - x[0]) + std::abs(S.rlc[i][1] - x[1]) + std::abs(S.rlc[i][2] - x[2]);
                    if (y < local_min) {
                        local_min = y;
                        local_min_which = i;
                    }
                }
                if (local_min < class_sum_cutoff) {
                    S.H_class[iRead] = local_min_which + 1;
                } else {
                    S.H_class[iRead] = 0;
                }
            }
        }
        iRead++;
        if ((nReads - 1) < iRead) {
            done_reads = true;
            read_wif_iRead = -1;
        } else {
            read_wif_iRead = S.R.wif0[iRead];
        }
    }
    if (at_least_one_read_has_changed) {
        for (int h = 0; h < S.nHaps; h++) {
            double* a = S.alphaHat_t[h].col(iGrid);
            const double* am = alphaHat_m.col(h);
            for (int kk = 0; kk < K; kk++) a[kk] = am[kk];
            alphaConst = 1 / accu2(am, K);
            S.c[h][iGrid] *= alphaConst;
            for (int kk = 0; kk < K; kk++) a[kk] *= alphaConst;
        }
    }
}

// ---------------------------------------------------------------- gibbs-nipt.cpp:1583-1621
void add_to_per_it_likelihoods(State& S, Mat& per_it_likelihoods, int i_gibbs_samplings, int iteration, int i_result_it,
                               int relabel, int& i_per_it_likelihoods) {
    int n = per_it_likelihoods.nr;
    if (i_per_it_likelihoods > n) return;
    if (i_per_it_likelihoods == n) return;  // reference would write out of bounds; never reached in supported modes
    double temp[7];
    calculate_likelihoods_values(S.c[0], S.c[1], S.c[2], S.H, S.nGrids, S.prior_probs, S.ff, temp);
    int r = i_per_it_likelihoods;
    per_it_likelihoods(r, 0) = 0 + 1;
    per_it_likelihoods(r, 1) = i_gibbs_samplings + 1;
    per_it_likelihoods(r, 2) = iteration + 1;
    per_it_likelihoods(r, 3) = i_result_it + 1;
    for (int j = 0; j < 7; j++) per_it_likelihoods(r, 4 + j) = temp[j];
    per_it_likelihoods(r, 11) = relabel;
    per_it_likelihoods(r, 12) = get_log_p_H_class(S.H_class, S.ff);
    i_per_it_likelihoods++;
}

// ---------------------------------------------------------------- gibbs-nipt.cpp:1629-1750
void gibbs_nipt_initialize(State& S, bool gibbs_initialize_iteratively) {
    const double prior = 1 / double(S.K);  // small_priorCurrent_m = 1 / Ksubset (quilt.R:756)
    for (int h = 0; h < 3; h++) S.eMatGrid_t[h].fill(1);
    if (!gibbs_initialize_iteratively) {
        for (int h = 0; h < S.nHaps; h++) make_eMatGrid_t(S.eMatGrid_t[h], S.eMatRead_t, S.H, S.R, h + 1);
        // diploid: hap 3 has no reads (H in {1,2}); its 1x1 buffer is untouched in the reference
    }
    if (gibbs_initialize_iteratively) {
        for (int h = 0; h < S.nHaps; h++) {
            S.alphaHat_t[h].fill(1);
            S.betaHat_t[h].fill(1);
            std::fill(S.c[h].begin(), S.c[h].end(), 1.0);
            run_forward_haploid(S.alphaHat_t[h], S.c[h], S.eMatGrid_t[h], prior, S.tm, true);
        }
    } else {
        for (int h = 0; h < S.nHaps; h++) {
            // rcpp_initialize_gibbs_forward_backward, gibbs-nipt.cpp:453-487
            run_forward_haploid(S.alphaHat_t[h], S.c[h], S.eMatGrid_t[h], prior, S.tm, false);
            double* bl = S.betaHat_t[h].col(S.nGrids - 1);
            for (int k = 0; k < S.K; k++) bl[k] = S.c[h][S.nGrids - 1];
            run_backward_haploid(S.betaHat_t[h], S.c[h], S.eMatGrid_t[h], prior, S.tm);
        }
    }
}

// ---------------------------------------------------------------- gibbs-nipt.cpp:1756-1956 (do_block_resampling = FALSE in production)
void gibbs_nipt_iterate(State& S, int iteration, const double* runif_reads, bool record_read_set, double class_sum_cutoff,
                        bool gibbs_initialize_iteratively, int first_read_for_gibbs_initialization, Mat& per_it_likelihoods,
                        int& i_per_it_likelihoods, int i_result_it) {
    const int K = S.K, nGrids = S.nGrids;
    bool done_reads = false;
    int iRead = -1;
    double minus_log_c_sum[3] = {0, 0, 0};  // only reported when return_p_store; value never read otherwise
    int read_wif_iRead = -1;
    int relabel = 1;
    Mat alphaHat_m(K, S.nHaps), betaHat_m(K, S.nHaps), ab_m(K, S.nHaps);
    double pC[3] = {1, 1, 1}, pA1[3] = {1, 1, 1}, pA2[3] = {1, 1, 1};
    const double prior = 1 / double(K);
    for (int iGrid = 0; iGrid < nGrids; iGrid++) {
        if (iGrid > 0) {
            for (int h = 0; h < S.nHaps; h++)
                alpha_forward_one_QUILT_faster(iGrid, K, S.alphaHat_t[h], S.tm, S.eMatGrid_t[h], S.c[h], minus_log_c_sum[h],
                                               S.grid_has_read, true);
        } else {
            for (int h = 0; h < S.nHaps; h++) reinitialize_in_iterations(S.alphaHat_t[h], S.c[h], prior, S.eMatGrid_t[h], K);
        }
        iRead++;
        if (!done_reads) {
            read_wif_iRead = S.R.wif0[iRead];
        } else {
            read_wif_iRead = -1;
        }
        if (read_wif_iRead == iGrid) {
            sample_reads_in_grid(S, iRead, iGrid, done_reads, read_wif_iRead, pC, pA1, pA2, alphaHat_m, betaHat_m, ab_m, runif_reads,
                                 iteration, record_read_set, class_sum_cutoff, gibbs_initialize_iteratively,
                                 first_read_for_gibbs_initialization);
        }
        iRead = iRead - 1;
    }
    for (int h = 0; h < S.nHaps; h++) {
        double* bl = S.betaHat_t[h].col(nGrids - 1);
        for (int k = 0; k < K; k++) bl[k] = S.c[h][nGrids - 1];
        run_backward_haploid_QUILT_faster(S.betaHat_t[h], S.c[h], S.eMatGrid_t[h], S.tm, S.grid_has_read);
    }
    add_to_per_it_likelihoods(S, per_it_likelihoods, 0, iteration, i_result_it, relabel, i_per_it_likelihoods);
}

// ---------------------------------------------------------------- gibbs-nipt-block.cpp:81-85
double simple_quantile(const Vec& x, double q) {
    int v = int(x.size() * q);
    std::vector<int> a(x.size());
    std::iota(a.begin(), a.end(), 0);
    std::stable_sort(a.begin(), a.end(), [&](int i, int j) { return x[i] < x[j]; });
    return x[a[v]];
}

// ---------------------------------------------------------------- copied-from-stitch.cpp:446-518
Vec make_smoothed_rate(const Vec& sigma_rate, const int32_t* L_grid, int nGrids, int shuffle_bin_radius) {
    Vec smoothed_rate(nGrids - 1, 0.0);
    for (int iGrid = 0; iGrid < (nGrids - 1); iGrid++) {
        int focal_point = (L_grid[iGrid] + L_grid[iGrid + 1]) / 2;
        int iGrid_left = iGrid;
        int bp_remaining = shuffle_bin_radius;
        int bp_prev = focal_point;
        double total_bp_added = 0;
        int bp_to_add;
        while ((0 < bp_remaining) & (0 <= iGrid_left)) {
            bp_to_add = (bp_prev - L_grid[iGrid_left]);
            if ((bp_remaining - bp_to_add) < 0) {
                bp_to_add = bp_remaining;
                bp_remaining = 0;
            } else {
                bp_remaining = bp_remaining - bp_to_add;
            }
            smoothed_rate[iGrid] = smoothed_rate[iGrid] + bp_to_add * sigma_rate[iGrid_left];
            total_bp_added += bp_to_add;
            bp_prev = L_grid[iGrid_left];
            iGrid_left = iGrid_left - 1;
        }
        int iGrid_right = iGrid + 1;
        bp_remaining = shuffle_bin_radius;
        bp_prev = focal_point;
        while ((0 < bp_remaining) & (iGrid_right < nGrids)) {
            bp_to_add = (L_grid[iGrid_right] - bp_prev);
            if ((bp_remaining - bp_to_add) < 0) {
                bp_to_add = bp_remaining;
                bp_remaining = 0;
            } else {
                bp_remaining = bp_remaining - bp_to_add;
            }
            smoothed_rate[iGrid] = smoothed_rate[iGrid] + bp_to_add * sigma_rate[iGrid_right - 1];
            total_bp_added += bp_to_add;
            bp_prev = L_grid[iGrid_right];
            iGrid_right = iGrid_right + 1;
        }
        smoothed_rate[iGrid] /= total_bp_added;
    }
    return smoothed_rate;
}

// ---------------------------------------------------------------- copied-from-stitch.cpp:522-567
int determine_where_to_stop(const Vec& smoothed_rate, const std::vector<uint8_t>& available, int snp_best, double thresh, int nGrids,
                            bool is_left) {
    double mult = is_left ? 1 : -1;
    int snp_consider = snp_best;
    double val_cur = smoothed_rate[snp_consider];
    double val_prev = smoothed_rate[snp_best];
    int snp_min = snp_consider;
    double val_min = smoothed_rate[snp_min];
    int c = 1;
    bool are_done = false;
    while (!are_done) {
        snp_consider = snp_consider + (-1) * mult;
        val_cur = smoothed_rate[snp_consider];
        if (5 <= c) val_prev = smoothed_rate[snp_consider + 5 * mult];
        c += 1;
        if (val_cur < val_min) {
            snp_min = snp_consider;
            val_min = val_cur;
        }
        if ((snp_consider <= 2) | ((nGrids - 3) <= snp_consider)) {
            are_done = true;
        } else if (available[snp_consider + (-1) * mult] == false) {
            are_done = true;
        } else if ((3 * val_min) < val_cur) {
            are_done = true;
        } else if ((val_cur < thresh) & (val_prev < val_cur)) {
            are_done = true;
        }
    }
    return snp_min;
}

// ---------------------------------------------------------------- gibbs-nipt-block.cpp:311-523
// returns blocked_snps [nSNPs]
IVec define_blocked_snps_using_gamma_on_the_fly(State& S, int nSNPs, const double* smooth_cm, int shuffle_bin_radius,
                                                const int32_t* L_grid, double block_gibbs_quantile_prob,
                                                bool use_smooth_cm_in_block_gibbs) {
    const int nGrids = S.nGrids, K = S.K;
    Vec rate2(nGrids - 1, 0.0);
    const int nh = (S.ff > 0) ? 3 : 2;
    for (int h = 0; h < nh; h++) {
        for (int iGrid = 0; iGrid < (nGrids - 2); iGrid++) {
            double d = S.tm(0, iGrid);
            const double* a = S.alphaHat_t[h].col(iGrid);
            const double* b = S.betaHat_t[h].col(iGrid + 1);
            const double* e = S.eMatGrid_t[h].col(iGrid + 1);
            rate2[iGrid] += 1 - d * accu2(K, [&](int k) { return a[k] * b[k] * e[k]; });
        }
    }
    Vec smoothed_rate = make_smoothed_rate(rate2, L_grid, nGrids, shuffle_bin_radius);
    if (use_smooth_cm_in_block_gibbs) {
        for (int iGrid = 0; iGrid < (nGrids - 2); iGrid++) rate2[iGrid] *= smooth_cm[iGrid];  // no effect on results (sic)
    }
    double break_thresh = 1;
    double d = simple_quantile(smoothed_rate, block_gibbs_quantile_prob);
    if (d < break_thresh) break_thresh = d;
    std::vector<uint8_t> available(nGrids - 1, 0);
    for (int i = 0; i < (nGrids - 1); i++) {
        if (smoothed_rate[i] < 0.01) available[i] = false;
        if (break_thresh < smoothed_rate[i]) available[i] = true;
    }
    IVec blocked_snps(nSNPs, 0);
    int nAvailable = 0;
    for (int i = 0; i < nGrids - 1; i++) nAvailable += available[i];
    if (nAvailable == 0) return blocked_snps;
    std::vector<int> best2(nGrids - 1);
    std::iota(best2.begin(), best2.end(), 0);
    std::stable_sort(best2.begin(), best2.end(), [&](int i, int j) { return smoothed_rate[i] > smoothed_rate[j]; });
    IVec to_keep;
    for (int iBest = 0; iBest < nAvailable; iBest++) {
        int snp_best = best2[iBest];
        if (available[snp_best]) {
            int a = std::max(snp_best - 1, 0);
            int b = std::min(snp_best + 1, nGrids - 1 - 1);
            double dd = 0;
            for (int j = a; j <= b; j++)
                if (available[j]) dd += 1;
            if (dd == 3) {
                int snp_left = determine_where_to_stop(smoothed_rate, available, snp_best, break_thresh, nGrids, true);
                int snp_right = determine_where_to_stop(smoothed_rate, available, snp_best, break_thresh, nGrids, false);
                for (int j = snp_left; j <= snp_right; j++) available[j] = false;
            } else {
                for (int j = a; j <= b; j++) available[j] = false;
            }
            to_keep.push_back(snp_best + 1);
        }
    }
    if (*std::min_element(to_keep.begin(), to_keep.end()) != 0) to_keep.push_back(0);
    if (*std::max_element(to_keep.begin(), to_keep.end()) != (nGrids - 1)) to_keep.push_back(nGrids - 1);
    std::sort(to_keep.begin(), to_keep.end());
    int n = (int)to_keep.size();
    IVec blocked_grid(nGrids, 0);
    for (int i = 0; i < (n - 1); i++) {
        int a = to_keep[i], b = to_keep[i + 1];
        for (int j = a; j <= b; j++) blocked_grid[j] = i;
    }
    for (int iSNP = 0; iSNP < nSNPs; iSNP++) blocked_snps[iSNP] = blocked_grid[iSNP / 32];
    return blocked_snps;
}

double ceiling_point5(double x) {
    if (double(int(x)) < x) return x + 0.5;
    return x;
}

struct Considers {
    IVec snp_start, snp_end, grid_start, grid_end, reads_start, reads_end, grid_where;
    int n_blocks = 0;
    bool ok = true;
};

// ---------------------------------------------------------------- gibbs-nipt-block.cpp:1307-1553
Considers make_gibbs_considers(const IVec& blocked_snps, const int32_t* wif0, int nReads, int nGrids) {
    Considers C;
    const int nSNPs = (int)blocked_snps.size();
    int n_blocks = blocked_snps[nSNPs - 1] + 1;
    for (int iSNP = 0; iSNP < (nSNPs - 1); iSNP++) {
        if ((blocked_snps[iSNP + 1] - blocked_snps[iSNP]) > 1) {
            C.ok = false;
            return C;
        }
    }
    IVec snp_start(n_blocks, 0), snp_end(n_blocks, 0);
    int start = 0;
    bool record = false;
    int iBlock = 0;
    for (int iSNP = 0; iSNP < nSNPs; iSNP++) {
        if (iSNP == (nSNPs - 1)) {
            record = true;
        } else if (blocked_snps[iSNP] < blocked_snps[iSNP + 1]) {
            record = true;
        } else {
            record = false;
        }
        if (record) {
            snp_start[iBlock] = start;
            snp_end[iBlock] = iSNP;
            start = iSNP + 1;
            iBlock = iBlock + 1;
        }
    }
    IVec grid_start(n_blocks), grid_end(n_blocks);
    for (iBlock = 0; iBlock < n_blocks; iBlock++) {
        grid_start[iBlock] = snp_start[iBlock] / 32;
        grid_end[iBlock] = snp_end[iBlock] / 32;
    }
    IVec blocked_grid(nGrids, 0);
    for (iBlock = 0; iBlock < n_blocks; iBlock++) {
        for (int i = grid_start[iBlock]; i <= grid_end[iBlock]; i++) blocked_grid[i] = iBlock;
    }
    IVec reads_start(n_blocks, -1), reads_end(n_blocks, -1);
    int previous_block_first_iRead = 0;
    int previous_grid = wif0[previous_block_first_iRead];
    int previous_block = blocked_grid[previous_grid];
    for (int this_iRead = 1; this_iRead < nReads; this_iRead++) {
        int this_grid = wif0[this_iRead];
        int this_block = blocked_grid[this_grid];
        if (this_iRead == (nReads - 1)) {
            reads_start[this_block] = previous_block_first_iRead;
            reads_end[this_block] = this_iRead;
        } else if (previous_block < this_block) {
            reads_start[previous_block] = previous_block_first_iRead;
            reads_end[previous_block] = this_iRead - 1;
            previous_block_first_iRead = this_iRead;
            previous_block = blocked_grid[wif0[this_iRead]];
            previous_grid = this_grid;
        }
    }
    (void)previous_grid;
    // do_removal = true
    {
        std::vector<uint8_t> remove(n_blocks);
        int n_to_remove = 0;
        for (iBlock = 0; iBlock < n_blocks; iBlock++) {
            remove[iBlock] = (reads_start[iBlock] == -1);
            n_to_remove += remove[iBlock];
        }
        if (n_to_remove > 0) {
            int n_new_blocks = n_blocks - n_to_remove;
            IVec w(n_to_remove);
            int a = 0;
            for (iBlock = 0; iBlock < n_blocks; iBlock++) {
                if (remove[iBlock]) {
                    w[a] = iBlock;
                    a += 1;
                }
            }
            int jBefore = 0;
            bool todo = false;
            for (int jNow = 0; jNow < n_to_remove; jNow++) {
                if (jNow == (n_to_remove - 1)) {
                    todo = true;
                } else {
                    if ((w[jNow + 1] - w[jNow]) == 1) {
                        todo = false;
                        jBefore -= 1;
                    } else {
                        todo = true;
                    }
                }
                if (todo) {
                    int s1 = w[jBefore];
                    int e1 = w[jNow];
                    double x = ceiling_point5(0.5 * double(grid_start[s1] + grid_end[e1]));
                    double y = ceiling_point5(0.5 * double(snp_start[s1] + snp_end[e1]));
                    if (s1 == 0) {
                        s1 = 1;
                        x = 0;
                        y = 0;
                    }
                    if (e1 == (n_blocks - 1)) {
                        e1 = e1 - 1;
                        x = grid_end[n_blocks - 1];
                        y = snp_end[n_blocks - 1];
                    }
                    grid_start[e1 + 1] = x;  // double -> int truncation, as Rcpp IntegerVector assignment
                    grid_end[s1 - 1] = x - 1;
                    snp_start[e1 + 1] = y;
                    snp_end[s1 - 1] = y - 1;
                    jBefore = jNow;
                }
                jBefore += 1;
            }
            IVec nrs(n_new_blocks), nre(n_new_blocks), ngs(n_new_blocks), nge(n_new_blocks), nss(n_new_blocks), nse(n_new_blocks);
            int i_prev_block = -1;
            for (iBlock = 0; iBlock < n_blocks; iBlock++) {
                if (!remove[iBlock]) {
                    i_prev_block += 1;
                    nrs[i_prev_block] = reads_start[iBlock];
                    nre[i_prev_block] = reads_end[iBlock];
                    ngs[i_prev_block] = grid_start[iBlock];
                    nge[i_prev_block] = grid_end[iBlock];
                    nss[i_prev_block] = snp_start[iBlock];
                    nse[i_prev_block] = snp_end[iBlock];
                }
            }
            reads_start = nrs;
            reads_end = nre;
            snp_start = nss;
            snp_end = nse;
            grid_start = ngs;
            grid_end = nge;
        }
    }
    n_blocks = (int)snp_end.size();
    IVec grid_where(nGrids, -1);
    for (iBlock = 0; iBlock < n_blocks; iBlock++) grid_where[grid_end[iBlock]] = iBlock;
    C.snp_start = snp_start;
    C.snp_end = snp_end;
    C.grid_start = grid_start;
    C.grid_end = grid_end;
    C.reads_start = reads_start;
    C.reads_end = reads_end;
    C.grid_where = grid_where;
    C.n_blocks = n_blocks;
    return C;
}

// rr (1-based permutations) and rr0 = rr - 1, gibbs-nipt-block.cpp:1760-1768
const int RR[6][3] = {{1, 2, 3}, {1, 3, 2}, {2, 1, 3}, {2, 3, 1}, {3, 1, 2}, {3, 2, 1}};
// rx, gibbs-nipt-block.cpp:766-773
const int RX[6][3] = {{1, 2, 3}, {1, 3, 2}, {2, 1, 3}, {3, 1, 2}, {2, 3, 1}, {3, 2, 1}};

struct Cube {  // arma::cube(n_rows, n_cols, n_slices), zero-filled
    int nr, nc, ns;
    std::vector<double> d;
    Cube(int r, int c, int s) : nr(r), nc(c), ns(s), d((size_t)r * c * s, 0.0) {}
    double* col(int slice, int c) { return d.data() + ((size_t)slice * nc + c) * nr; }
    double& operator()(int i, int j, int s) { return d[((size_t)s * nc + j) * nr + i]; }
};

// ---------------------------------------------------------------- gibbs-nipt-block.cpp:1122-1253 (block_approach = 6)
void gibbs_block_forward_one(int iGrid, double ff, Cube& alphaStore, Cube& log_cStore, Mat& eMatGridLocal, const Trans& tm, int K) {
    const double one_over_K = 1 / double(K);
    const double prior = 1 / double(K);
    if (iGrid == 0) {
        for (int ir = 0; ir < 6; ir++) {
            for (int i = 0; i < 3; i++) {
                int h = RR[ir][i] - 1;
                double* as = alphaStore.col(ir, h);
                const double* el = eMatGridLocal.col(i);
                for (int k = 0; k < K; k++) as[k] = prior * el[k];
                double d = 1 / accu2(as, K);
                log_cStore(iGrid, h, ir) = std::log(d);
                for (int k = 0; k < K; k++) as[k] = d * as[k];
            }
        }
    } else {
        const double t0 = tm(0, iGrid - 1), t1 = tm(1, iGrid - 1);
        for (int ir = 0; ir < 6; ir++) {
            if ((ff > 0) | ((ff == 0) & ((ir == 0) | (ir == 2)))) {
                for (int i = 0; i < 3; i++) {
                    int h = RR[ir][i] - 1;
                    double* as = alphaStore.col(ir, h);
                    const double* el = eMatGridLocal.col(i);
                    const double jump = t1 * one_over_K;
                    for (int k = 0; k < K; k++) as[k] = el[k] * (t0 * as[k] + jump);
                    double d = 1 / accu2(as, K);
                    log_cStore(iGrid, h, ir) = std::log(d);
                    for (int k = 0; k < K; k++) as[k] = d * as[k];
                }
            }
        }
    }
}

// ---------------------------------------------------------------- gibbs-nipt-block.cpp:590-949 (block_approach = 6)
void consider_block_relabelling(State& S, int iBlock, const double* runif_block, double sum_H[3], const double log_prior_probs[3],
                                double logC_before[3], double logC_after[3], Mat& eMatGridLocal, Mat& betaHatLocal, int iGrid,
                                int grid_start_0_based, int grid_end_0_based, int read_start_0_based, int read_end_0_based,
                                Cube& log_cStore, Cube& alphaStore, double& ever_changed, Mat* block_results) {
    const int K = S.K, nReads = S.nReads;
    const bool sample_is_diploid = S.sample_is_diploid;
    const double ff = S.ff;
    const double prior = 1 / double(K);
    for (int h = 0; h < S.nHaps; h++) {
        const double* b = S.betaHat_t[h].col(iGrid);
        double* bl = betaHatLocal.col(h);
        for (int k = 0; k < K; k++) bl[k] = b[k];
    }
    double choice_log_probs_Pm[6][3];
    double choice_log_probs_P[6] = {0, 0, 0, 0, 0, 0};
    for (int ir = 0; ir < 6; ir++) {
        for (int i = 0; i < 3; i++) {
            double logC_inside = 0;
            for (int iGrid2 = grid_start_0_based; iGrid2 <= grid_end_0_based; iGrid2++) logC_inside += log_cStore(iGrid2, i, ir);
            const double* as = alphaStore.col(ir, i);
            const double* bl = betaHatLocal.col(i);
            choice_log_probs_Pm[ir][i] =
                std::log(accu2(K, [&](int k) { return as[k] * bl[k]; })) + -logC_before[i] + -logC_inside + -logC_after[i];
            choice_log_probs_P[ir] += choice_log_probs_Pm[ir][i];
        }
    }
    // rcpp_calculate_block_read_label_probabilities_using_H_class, block.cpp:251-279
    double choice_log_probs_H[6];
    {
        int ns[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int iRead = read_start_0_based; iRead <= read_end_0_based; iRead++) ns[S.H_class[iRead]]++;
        for (int ir = 0; ir < 6; ir++) {
            choice_log_probs_H[ir] = get_log_p_H_class2(ns[RR[ir][0]], ns[RR[ir][1]], ns[RR[ir][2]], ns[7 - RR[ir][2]],
                                                        ns[7 - RR[ir][1]], ns[7 - RR[ir][0]], ff);
        }
    }
    double choice_log_probs[6], choice_probs[6];
    for (int ir = 0; ir < 6; ir++) choice_log_probs[ir] = choice_log_probs_H[ir] + choice_log_probs_P[ir];
    double a = -(rcpp_max(choice_log_probs, 6));
    for (int ir = 0; ir < 6; ir++) choice_log_probs[ir] += a;
    for (int ir = 0; ir < 6; ir++) {
        if (choice_log_probs[ir] < (-100)) choice_log_probs[ir] = -100;
        choice_probs[ir] = std::exp(choice_log_probs[ir]);
    }
    if (ff == 0) {
        choice_probs[1] = 0;
        choice_probs[3] = 0;
        choice_probs[4] = 0;
        choice_probs[5] = 0;
    }
    double ssum = 0;
    for (int ir = 0; ir < 6; ir++) ssum += choice_probs[ir];  // Rcpp sugar sum: sequential
    double d = (1 / ssum);
    for (int ir = 0; ir < 6; ir++) choice_probs[ir] *= d;
    double chance = runif_block[iBlock];
    int ir_chosen = 0;
    double cumsum_probs[6] = {0, 0, 0, 0, 0, 0};
    cumsum_probs[0] = choice_probs[0];
    for (int ir = 1; ir < 6; ir++) cumsum_probs[ir] += choice_probs[ir] + cumsum_probs[ir - 1];
    for (int ir = 5; ir >= 0; ir--) {
        if (chance < cumsum_probs[ir]) ir_chosen = ir;
    }
    int zero_based_swap[8];
    zero_based_swap[0] = 0;
    zero_based_swap[1] = RX[ir_chosen][0];
    zero_based_swap[2] = RX[ir_chosen][1];
    zero_based_swap[3] = RX[ir_chosen][2];
    zero_based_swap[4] = 7 - RX[ir_chosen][2];
    zero_based_swap[5] = 7 - RX[ir_chosen][1];
    zero_based_swap[6] = 7 - RX[ir_chosen][0];
    zero_based_swap[7] = 7;
    if (block_results) {
        Mat& br = *block_results;
        int ibr = 2 * iBlock;
        br(ibr, 0) = iBlock;
        br(ibr, 1) = 0;
        for (int ir = 0; ir < 6; ir++) br(ibr, 2 + ir) = choice_probs[ir];
        br(ibr, 8) = ir_chosen + 1;
        br(ibr, 9) = choice_log_probs_Pm[ir_chosen][0];
        br(ibr, 10) = choice_log_probs_Pm[ir_chosen][1];
        br(ibr, 11) = choice_log_probs_Pm[ir_chosen][2];
        for (int i = 0; i < 3; i++) br(ibr, 12) += choice_log_probs_Pm[ir_chosen][i];
    }
    if (!((ever_changed == 1) | (ir_chosen != 0))) {
        // no change warranted
    } else {
        ever_changed = 1;
        int iRead = read_start_0_based;
        int wif_read = S.R.wif0[iRead];
        int h = -1;
        for (int iGrid2 = grid_start_0_based; iGrid2 <= grid_end_0_based; iGrid2++) {
            eMatGridLocal.fill(1);
            while ((iRead <= (nReads - 1)) & (wif_read < iGrid2)) {
                iRead += 1;
                if (iRead < (nReads - 1)) wif_read = S.R.wif0[iRead];
            }
            while ((iRead <= (nReads - 1)) & (wif_read == iGrid2)) {
                h = zero_based_swap[S.H[iRead]] - 1;
                double* el = eMatGridLocal.col(h);
                const double* er = S.eMatRead_t.col(iRead);
                for (int k = 0; k < K; k++) el[k] *= er[k];
                iRead += 1;
                if (iRead <= (nReads - 1)) wif_read = S.R.wif0[iRead];
            }
            for (int hh = 0; hh < S.nHaps; hh++) {
                double* g = S.eMatGrid_t[hh].col(iGrid2);
                const double* el = eMatGridLocal.col(hh);
                for (int k = 0; k < K; k++) g[k] = el[k];
            }
            if (iGrid2 == 0) {
                for (int hh = 0; hh < S.nHaps; hh++) {
                    double* al = S.alphaHat_t[hh].col(0);
                    const double* g = S.eMatGrid_t[hh].col(0);
                    for (int k = 0; k < K; k++) al[k] = prior * g[k];
                }
            } else {
                const double t0 = S.tm(0, iGrid2 - 1), t1 = S.tm(1, iGrid2 - 1);
                for (int hh = 0; hh < S.nHaps; hh++) {
                    double* al = S.alphaHat_t[hh].col(iGrid2);
                    const double* ap = S.alphaHat_t[hh].col(iGrid2 - 1);
                    const double* g = S.eMatGrid_t[hh].col(iGrid2);
                    for (int k = 0; k < K; k++) al[k] = g[k] * (t0 * ap[k] + t1 * prior);
                }
            }
            for (int hh = 0; hh < S.nHaps; hh++) {
                double* al = S.alphaHat_t[hh].col(iGrid2);
                S.c[hh][iGrid2] = 1 / accu2(al, K);
                for (int k = 0; k < K; k++) al[k] *= S.c[hh][iGrid2];
            }
        }
        for (int iRead2 = read_start_0_based; iRead2 <= read_end_0_based; iRead2++) {
            int lost = S.H[iRead2] - 1;
            int gained = zero_based_swap[S.H[iRead2]] - 1;
            S.H_class[iRead2] = zero_based_swap[S.H_class[iRead2]];
            S.H[iRead2] = gained + 1;
            sum_H[gained] += 1.0;
            sum_H[lost] -= 1.0;
        }
    }
    (void)sample_is_diploid;
    (void)log_prior_probs;
}

// ---------------------------------------------------------------- gibbs-nipt-block.cpp:1257-1292
void reset_local_variables(State& S, int iGrid, Mat& alphaHatLocal, Cube& alphaStore, Cube& log_cStore) {
    const int K = S.K;
    for (int h = 0; h < 3; h++) {
        double* al = alphaHatLocal.col(h);
        if (h < S.nHaps) {
            const double* a = S.alphaHat_t[h].col(iGrid);
            for (int k = 0; k < K; k++) al[k] = a[k];
        } else {
            for (int k = 0; k < K; k++) al[k] = 0;
        }
    }
    double cLocal[3] = {S.c[0][iGrid], S.c[1][iGrid], S.c[2][iGrid]};
    for (int ir = 0; ir < 6; ir++) {
        for (int i = 0; i < 3; i++) {
            double* as = alphaStore.col(ir, i);
            const double* al = alphaHatLocal.col(i);
            for (int k = 0; k < K; k++) as[k] = al[k];
            log_cStore(iGrid, i, ir) = std::log(cLocal[i]);
        }
    }
}

// R's revsort for tiny n is not needed: Rcpp::sample(1:3, 1, false, probs) (sugar/functions/sample.h,
// SampleReplace path because size < 2): normalise p, sort descending carrying the index, cumulate,
// pick the first j with rU <= p[j].  Ties keep ascending index order here (three-element case).
int rcpp_sample_1_of_3(const double probs_in[3], double rU) {
    double p[3];
    int perm[3] = {1, 2, 3};
    double s = probs_in[0] + probs_in[1] + probs_in[2];
    for (int i = 0; i < 3; i++) p[i] = probs_in[i] / s;
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2 - i; j++)
            if (p[j] < p[j + 1]) {
                std::swap(p[j], p[j + 1]);
                std::swap(perm[j], perm[j + 1]);
            }
    for (int i = 1; i < 3; i++) p[i] += p[i - 1];
    int j;
    for (j = 0; j < 2; j++)
        if (rU <= p[j]) break;
    return perm[j];
}

// ---------------------------------------------------------------- gibbs-nipt-block.cpp:213-246
void sample_H_using_H_class(IVec& H_class, IVec& H, double ff, const double* runif_H_class, int& n_used) {
    const double probs07[3] = {0.5, 0.5 - ff * 0.5, ff * 0.5};
    const double probs4[3] = {0.5, 0.5 - 0.5 * ff, 0};
    const double probs5[3] = {0.5, 0, 0.5 * ff};
    const double probs6[3] = {0, 0.5 - ff * 0.5, ff * 0.5};
    int i_temp = 0;  // i_temp persists across reads in the reference (IntegerVector i_temp(1))
    for (size_t iRead = 0; iRead < H.size(); iRead++) {
        int hc = H_class[iRead];
        if (hc == 0 || hc == 7) {
            i_temp = rcpp_sample_1_of_3(probs07, runif_H_class[n_used++]);
        } else if (hc == 1) {
            i_temp = 1;
        } else if (hc == 2) {
            i_temp = 2;
        } else if (hc == 3) {
            i_temp = 3;
        } else if (hc == 4) {
            i_temp = rcpp_sample_1_of_3(probs4, runif_H_class[n_used++]);
        } else if (hc == 5) {
            i_temp = rcpp_sample_1_of_3(probs5, runif_H_class[n_used++]);
        } else if (hc == 6) {
            i_temp = rcpp_sample_1_of_3(probs6, runif_H_class[n_used++]);
        }
        H[iRead] = i_temp;
    }
}

// ---------------------------------------------------------------- gibbs-nipt-block.cpp:1636-1967
// block_approach = 6, consider_total_relabelling = false, resample_H_using_H_class = true (defaults, not overridden at gibbs-nipt.cpp:3025)
// returns the number of uniforms rcpp_sample_H_using_H_class consumed
int block_gibbs_resampler(State& S, const IVec& blocked_snps, const double* runif_block, const double* runif_H_class) {
    const int K = S.K, nGrids = S.nGrids, nReads = S.nReads;
    const double ff = S.ff;
    const double prior = 1 / double(K);
    double log_prior_probs[3] = {std::log(0.5), std::log((1 - ff) / 2), std::log((ff / 2))};
    Considers C = make_gibbs_considers(blocked_snps, S.R.wif0, nReads, nGrids);
    if (!C.ok) return 0;  // reference: out["..."] on an empty list would throw; unreachable for valid blocked_snps
    int n_used_H = 0;
    const int n_blocks = C.n_blocks;
    double logC_before[3] = {0, 0, 0};
    double logC_after[3];
    for (int h = 0; h < 3; h++) logC_after[h] = accu2(nGrids, [&](int g) { return std::log(S.c[h][g]); });
    int iBlock = 0;
    double ever_changed = 0;
    double sum_H[3] = {0, 0, 0};
    for (int iRead = 0; iRead < nReads; iRead++) sum_H[S.H[iRead] - 1] += 1;
    Cube alphaStore(K, 3, 6);
    Mat alphaHatLocal(K, 3), betaHatLocal(K, 3), eMatGridLocal(K, 3);
    Cube log_cStore(nGrids, 3, 6);
    for (int iGrid = 0; iGrid < nGrids; iGrid++) {
        for (int h = 0; h < S.nHaps; h++) {
            const double* g = S.eMatGrid_t[h].col(iGrid);
            double* el = eMatGridLocal.col(h);
            for (int k = 0; k < K; k++) el[k] = g[k];
        }
        gibbs_block_forward_one(iGrid, ff, alphaStore, log_cStore, eMatGridLocal, S.tm, K);
        if ((-1) < C.grid_where[iGrid]) {
            iBlock = C.grid_where[iGrid];
            int grid_start_0_based = C.grid_start[iBlock];
            int grid_end_0_based = C.grid_end[iBlock];
            int read_start_0_based = C.reads_start[iBlock];
            int read_end_0_based = C.reads_end[iBlock];
            consider_block_relabelling(S, iBlock, runif_block, sum_H, log_prior_probs, logC_before, logC_after, eMatGridLocal,
                                       betaHatLocal, iGrid, grid_start_0_based, grid_end_0_based, read_start_0_based,
                                       read_end_0_based, log_cStore, alphaStore, ever_changed, nullptr);
            if ((iBlock + 1) < n_blocks) reset_local_variables(S, iGrid, alphaHatLocal, alphaStore, log_cStore);
            for (int iGrid2 = grid_start_0_based; iGrid2 <= grid_end_0_based; iGrid2++) {
                for (int h = 0; h < 3; h++) logC_before[h] += std::log(S.c[h][iGrid2]);
            }
        }
        for (int h = 0; h < 3; h++) logC_after[h] -= std::log(S.c[h][iGrid]);
    }
    if ((ff > 0) && true && true) {
        int n_used = 0;
        sample_H_using_H_class(S.H_class, S.H, ff, runif_H_class, n_used);
        n_used_H = n_used;
        for (int h = 0; h < 3; h++) S.eMatGrid_t[h].fill(1);
        for (int h = 0; h < 3; h++) make_eMatGrid_t(S.eMatGrid_t[h], S.eMatRead_t, S.H, S.R, h + 1);
        for (int h = 0; h < 3; h++) {
            S.alphaHat_t[h].fill(1);
            S.betaHat_t[h].fill(1);
            std::fill(S.c[h].begin(), S.c[h].end(), 1.0);
            run_forward_haploid(S.alphaHat_t[h], S.c[h], S.eMatGrid_t[h], prior, S.tm, false);
        }
    }
    for (int h = 0; h < S.nHaps; h++) {
        double* bl = S.betaHat_t[h].col(nGrids - 1);
        for (int k = 0; k < K; k++) bl[k] = S.c[h][nGrids - 1];
    }
    // both the generic and the fast backward run; the second overwrites the first (block.cpp:1950-1958)
    for (int h = 0; h < S.nHaps; h++) run_backward_haploid(S.betaHat_t[h], S.c[h], S.eMatGrid_t[h], prior, S.tm);
    for (int h = 0; h < S.nHaps; h++)
        run_backward_haploid_QUILT_faster(S.betaHat_t[h], S.c[h], S.eMatGrid_t[h], S.tm, S.grid_has_read);
    return n_used_H;
}

// ---------------------------------------------------------------- gibbs-nipt-block.cpp:1975-2355 (ff == 0 branch; shard is diploid-only, functions.R:2552-2556)
int shard_block_gibbs_resampler(State& S, const IVec& blocked_snps, bool shard_check_every_pair, const double* runif_block) {
    const int K = S.K, nGrids = S.nGrids, nReads = S.nReads;
    const double ff = S.ff;
    if (ff > 0) return QUILT_ERR_UNSUPPORTED;  // the reference never calls the shard pass with ff > 0
    const double prior = 1 / double(K);
    Considers C;
    int n_blocks;
    if (!shard_check_every_pair) {
        C = make_gibbs_considers(blocked_snps, S.R.wif0, nReads, nGrids);
        n_blocks = C.n_blocks;
    } else {
        n_blocks = nGrids;
    }
    (void)n_blocks;
    double minus_log_c_sum = 0;
    double minus_log_c1_sum = 0, minus_log_c2_sum = 0;
    double minus_log_original_c1_sum = 0, minus_log_original_c2_sum = 0;
    double original_c1_this_grid, original_c2_this_grid;
    double pA1, pA2, pB1, pB2;
    for (int iGrid2 = 0; iGrid2 < nGrids; iGrid2++) {
        minus_log_original_c1_sum -= std::log(S.c[0][iGrid2]);
        minus_log_original_c2_sum -= std::log(S.c[1][iGrid2]);
    }
    bool in_flip_mode = false;
    int iRead = 0;
    int iGridConsider;
    bool done_reads = false;
    double calculated_difference, probs1, probs2, probs_sum;
    Mat& a1 = S.alphaHat_t[0];
    Mat& a2 = S.alphaHat_t[1];
    Mat& b1 = S.betaHat_t[0];
    Mat& b2 = S.betaHat_t[1];
    Mat& e1 = S.eMatGrid_t[0];
    Mat& e2 = S.eMatGrid_t[1];
    Vec& c1 = S.c[0];
    Vec& c2 = S.c[1];
    for (int iGrid = 0; iGrid < nGrids; iGrid++) {
        original_c1_this_grid = c1[iGrid];
        original_c2_this_grid = c2[iGrid];
        if (iGrid == 0) {
            double* x1 = a1.col(0);
            const double* g1 = e1.col(0);
            for (int k = 0; k < K; k++) x1[k] = prior * g1[k];
            c1[0] = 1 / accu2(x1, K);
            for (int k = 0; k < K; k++) x1[k] *= c1[0];
            double* x2 = a2.col(0);
            const double* g2 = e2.col(0);
            for (int k = 0; k < K; k++) x2[k] = prior * g2[k];
            c2[0] = 1 / accu2(x2, K);
            for (int k = 0; k < K; k++) x2[k] *= c2[0];
        } else {
            if (ff == 0 && in_flip_mode) {
                double* g1 = e1.col(iGrid);
                double* g2 = e2.col(iGrid);
                for (int k = 0; k < K; k++) std::swap(g1[k], g2[k]);
            }
            alpha_forward_one(iGrid, K, a1, S.tm, e1, prior, c1, minus_log_c_sum, true);
            alpha_forward_one(iGrid, K, a2, S.tm, e2, prior, c2, minus_log_c_sum, true);
        }
        minus_log_c1_sum -= std::log(c1[iGrid]);
        minus_log_c2_sum -= std::log(c2[iGrid]);
        done_reads = false;
        while (!done_reads) {
            if (iRead > (nReads - 1)) {
                done_reads = true;
            } else {
                if (S.R.wif0[iRead] == iGrid) {
                    if (ff == 0 && in_flip_mode) S.H[iRead] = 3 - S.H[iRead];
                    iRead++;
                }
                if (iRead > (nReads - 1)) {
                    done_reads = true;
                } else {
                    if (S.R.wif0[iRead] > iGrid) done_reads = true;
                }
            }
        }
        bool check = false;
        if (!shard_check_every_pair) {
            iGridConsider = C.grid_where[iGrid];
            if ((-1 < iGridConsider) & (iGridConsider < (C.n_blocks - 1))) check = true;
        } else {
            if (iGrid < (nGrids - 1)) check = true;
            iGridConsider = iGrid;
        }
        if (check) {
            const double* x1 = a1.col(iGrid);
            const double* x2 = a2.col(iGrid);
            const double* y1 = b1.col(iGrid);
            const double* y2 = b2.col(iGrid);
            pA1 = minus_log_c1_sum + minus_log_original_c1_sum + std::log(accu2(K, [&](int k) { return x1[k] * y1[k]; }));
            pA2 = minus_log_c2_sum + minus_log_original_c2_sum + std::log(accu2(K, [&](int k) { return x2[k] * y2[k]; }));
            pB1 = minus_log_c2_sum + minus_log_original_c1_sum + std::log(accu2(K, [&](int k) { return x2[k] * y1[k]; }));
            pB2 = minus_log_c1_sum + minus_log_original_c2_sum + std::log(accu2(K, [&](int k) { return x1[k] * y2[k]; }));
            calculated_difference = pB1 + pB2 - pA1 - pA2;
            probs1 = 1;
            probs2 = std::exp(calculated_difference);
            probs_sum = probs1 + probs2;
            probs1 /= probs_sum;
            probs2 /= probs_sum;
            in_flip_mode = runif_block[iGridConsider] > probs1;
        }
        minus_log_original_c1_sum += std::log(original_c1_this_grid);
        minus_log_original_c2_sum += std::log(original_c2_this_grid);
    }
    for (int h = 0; h < 2; h++) {
        double* bl = S.betaHat_t[h].col(nGrids - 1);
        for (int k = 0; k < K; k++) bl[k] = S.c[h][nGrids - 1];
        run_backward_haploid(S.betaHat_t[h], S.c[h], S.eMatGrid_t[h], prior, S.tm);
    }
    return QUILT_OK;
}

// ---------------------------------------------------------------- gibbs-small.cpp:472-635 (calculate_gamma_on_the_fly = TRUE)
void calculate_genProbs_and_hapProbs_using_binary_objects(State& S, const QuiltPanel* p, const int32_t* which, Mat& genProbsM_t,
                                                           Mat& genProbsF_t, Mat& hapProbs_t) {
    const int K = S.K, nGrids = S.nGrids;
    const int nSNPs = genProbsM_t.nc;
    const double ref_error = p->ref_error;
    const double ref_one_minus_error = 1 - ref_error;
    double g[3][32], hh[3][32];
    Vec gammaLocal[3] = {Vec(K, 0.0), Vec(K, 0.0), Vec(K, 0.0)};
    for (int iGrid = 0; iGrid < nGrids; iGrid++) {
        int s = 32 * iGrid;
        int e = 32 * (iGrid + 1) - 1;
        if (e > (nSNPs - 1)) e = nSNPs - 1;
        int nSNPsLocal = e - s + 1;
        for (int h = 0; h < 3; h++)
            for (int b = 0; b < 32; b++) g[h][b] = hh[h][b] = 0;
        for (int h = 0; h < S.nHaps; h++) {
            double x = 1 / S.c[h][iGrid];
            const double* a = S.alphaHat_t[h].col(iGrid);
            const double* bt = S.betaHat_t[h].col(iGrid);
            for (int k = 0; k < K; k++) gammaLocal[h][k] = (a[k] * bt[k]) * x;
        }
        for (int k = 0; k < K; k++) {
            double gk0 = gammaLocal[0][k], gk1 = gammaLocal[1][k], gk2 = gammaLocal[2][k];
            uint32_t tmp = (uint32_t)panel_word(p, which[k] - 1, iGrid);
            for (int b = 0; b < nSNPsLocal; b++, tmp >>= 1) {
                if ((tmp & 0x1) == 0) {
                    hh[0][b] += gk0;
                    hh[1][b] += gk1;
                    hh[2][b] += gk2;
                } else {
                    g[0][b] += gk0;
                    g[1][b] += gk1;
                    g[2][b] += gk2;
                }
            }
        }
        for (int b = 0; b < nSNPsLocal; b++)
            for (int h = 0; h < 3; h++) g[h][b] = g[h][b] * ref_one_minus_error + hh[h][b] * ref_error;
        for (int b = 0; b < nSNPsLocal; b++) {
            double g0 = g[0][b], g1 = g[1][b], g2 = g[2][b];
            genProbsM_t(0, s + b) = (1 - g0) * (1 - g1);
            genProbsM_t(1, s + b) = (g0 * (1 - g1) + (1 - g0) * g1);
            genProbsM_t(2, s + b) = g0 * g1;
            genProbsF_t(0, s + b) = (1 - g0) * (1 - g2);
            genProbsF_t(1, s + b) = (g0 * (1 - g2) + (1 - g0) * g2);
            genProbsF_t(2, s + b) = g0 * g2;
            hapProbs_t(0, s + b) = g0;
            hapProbs_t(1, s + b) = g1;
            hapProbs_t(2, s + b) = g2;
        }
    }
}

// ---------------------------------------------------------------- gibbs-small.cpp:711-867
void calculate_genProbs_and_hapProbs_final_rare_common(State& S, const QuiltPanel* p, const int32_t* which, Mat& genProbsM_t,
                                                        Mat& genProbsF_t, Mat& hapProbs_t,
                                                        const std::vector<std::vector<int>>& rare_per_snp) {
    const int K = S.K;
    const int nSNPs = genProbsM_t.nc;
    const double ref_error = p->ref_error;
    const bool sample_is_diploid = S.sample_is_diploid;
    Vec gammaLocal[3] = {Vec(K, 0.0), Vec(K, 0.0), Vec(K, 0.0)};
    int iFullGrid_prev = -1;
    double one_minus_2_times_ref_error = 1 - 2 * ref_error;
    for (int iFullSNP = 0; iFullSNP < nSNPs; iFullSNP++) {
        int iFullGrid = iFullSNP / 32;
        if (iFullGrid != iFullGrid_prev) {
            for (int h = 0; h < S.nHaps; h++) {
                double x = 1 / S.c[h][iFullGrid];
                const double* a = S.alphaHat_t[h].col(iFullGrid);
                const double* bt = S.betaHat_t[h].col(iFullGrid);
                for (int k = 0; k < K; k++) gammaLocal[h][k] = (a[k] * bt[k]) * x;
            }
            iFullGrid_prev = iFullGrid;
        }
        if (p->snp_is_common[iFullSNP]) {
            int common_snp = p->common_snp_index[iFullSNP] - 1;
            int common_grid = common_snp / 32;
            for (int k = 0; k < K; k++) {
                int kk = p->hapMatcherR[(size_t)common_grid * p->K_full + (which[k] - 1)];
                double d;
                if (kk > 0) {
                    d = p->distinctHapsIE[(size_t)common_snp * p->nMaxDH + (kk - 1)];
                } else {
                    int bvtd = simple_binary_matrix_search(which[k] - 1, p->eMatDH_special_matrix, p->n_special,
                                                           p->eMatDH_special_matrix_helper[common_grid],
                                                           p->eMatDH_special_matrix_helper[(size_t)p->nGrids + common_grid]);
                    uint32_t tmp = (uint32_t)bvtd;
                    if (((tmp >> (common_snp % 32)) & 0x1) == 1) {
                        d = 1 - ref_error;
                    } else {
                        d = ref_error;
                    }
                }
                hapProbs_t(0, iFullSNP) += gammaLocal[0][k] * d;
                hapProbs_t(1, iFullSNP) += gammaLocal[1][k] * d;
                if (!sample_is_diploid) hapProbs_t(2, iFullSNP) += gammaLocal[2][k] * d;
            }
        } else {
            const std::vector<int>& k_with_alt = rare_per_snp[iFullSNP];
            if (k_with_alt.empty()) {
                hapProbs_t(0, iFullSNP) = ref_error;
                hapProbs_t(1, iFullSNP) = ref_error;
                if (!sample_is_diploid) hapProbs_t(2, iFullSNP) = ref_error;
            } else {
                for (int k = 0; k < K; k++) {
                    hapProbs_t(0, iFullSNP) += gammaLocal[0][k] * ref_error;
                    hapProbs_t(1, iFullSNP) += gammaLocal[1][k] * ref_error;
                    if (!sample_is_diploid) hapProbs_t(2, iFullSNP) += gammaLocal[2][k] * ref_error;
                }
                for (size_t ik = 0; ik < k_with_alt.size(); ik++) {
                    int k = k_with_alt[ik] - 1;
                    hapProbs_t(0, iFullSNP) += gammaLocal[0][k] * one_minus_2_times_ref_error;
                    hapProbs_t(1, iFullSNP) += gammaLocal[1][k] * one_minus_2_times_ref_error;
                    if (!sample_is_diploid) hapProbs_t(2, iFullSNP) += gammaLocal[2][k] * one_minus_2_times_ref_error;
                }
            }
        }
        double h1 = hapProbs_t(0, iFullSNP);
        double h2 = hapProbs_t(1, iFullSNP);
        genProbsM_t(0, iFullSNP) = (1 - h1) * (1 - h2);
        genProbsM_t(1, iFullSNP) = h1 * (1 - h2) + h2 * (1 - h1);
        genProbsM_t(2, iFullSNP) = h1 * h2;
        if (!sample_is_diploid) {
            h1 = hapProbs_t(0, iFullSNP);
            h2 = hapProbs_t(2, iFullSNP);
            genProbsF_t(0, iFullSNP) = (1 - h1) * (1 - h2);
            genProbsF_t(1, iFullSNP) = h1 * (1 - h2) + h2 * (1 - h1);
            genProbsF_t(2, iFullSNP) = h1 * h2;
        }
    }
}

void setup_state(State& S, const QuiltGibbsArgs* a) {
    S.K = a->K;
    S.nGrids = a->nGrids;
    S.nReads = a->reads.nReads;
    S.sample_is_diploid = (a->flags & QUILT_F_SAMPLE_IS_DIPLOID) != 0;
    S.nHaps = S.sample_is_diploid ? 2 : 3;
    S.ff = a->ff;
    S.tm.t = a->transMatRate_tc_H;
    S.R = Reads{a->reads.nReads, a->reads.offsets, a->reads.u, a->reads.bq, a->reads.wif0};
    S.prior_probs[0] = 0.5;
    S.prior_probs[1] = (1 - a->ff) * 0.5;
    S.prior_probs[2] = a->ff * 0.5;
    const double* pp = S.prior_probs;
    double rlc[7][3] = {{1, 0, 0},
                        {0, 1, 0},
                        {0, 0, 1},
                        {pp[0] / (pp[0] + pp[1]), pp[1] / (pp[0] + pp[1]), 0},
                        {pp[0] / (pp[0] + pp[2]), 0, pp[2] / (pp[0] + pp[2])},
                        {0, pp[1] / (pp[1] + pp[2]), pp[2] / (pp[1] + pp[2])},
                        {pp[0], pp[1], pp[2]}};
    std::memcpy(S.rlc, rlc, sizeof(rlc));
    S.grid_has_read.assign(S.nGrids, 0);
    for (int r = 0; r < S.nReads; r++) S.grid_has_read[a->reads.wif0[r]] = 1;  // functions.R:314-316
}

void build_eMatRead(State& S, const QuiltGibbsArgs* a, std::vector<std::vector<int>>& rare_per_snp) {
    S.eMatRead_t = Mat(S.K, S.nReads, 1.0);
    const bool rescale = (a->flags & QUILT_F_RESCALE_EMATREAD) != 0;
    if (a->flags & QUILT_F_MAKE_EMATREAD_RARE_COMMON) {
        build_rare_per_snp(a->panel, a->K, a->which_haps_to_use, rare_per_snp);
        make_eMatRead_t_rare_common(S.eMatRead_t, S.R, a->panel, a->which_haps_to_use, rescale, a->Jmax, a->maxDifferenceBetweenReads,
                                    rare_per_snp);
    } else {
        make_eMatRead_t_using_objects(S.eMatRead_t, S.R, a->panel, a->which_haps_to_use, rescale, a->Jmax,
                                      a->maxDifferenceBetweenReads);
    }
    S.number_of_non_1_reads.assign(S.nReads, 0);
    S.read_category.assign(S.nReads, 0);
    S.indices_of_non_1_reads.assign((size_t)S.K * S.nReads, 0);
    evaluate_read_variability(S.eMatRead_t, S.number_of_non_1_reads, S.indices_of_non_1_reads, S.read_category);
    if (a->flags & QUILT_F_FORCE_RESET_READ_CATEGORY_0) {
        for (int r = 0; r < S.nReads; r++)
            if (S.read_category[r] != 1) S.read_category[r] = 0;
    }
    if (a->flags & QUILT_F_DISABLE_READ_CATEGORY_USAGE) std::fill(S.read_category.begin(), S.read_category.end(), 0);
}

}  // namespace

// ================================================================ exported entry points
extern "C" {

// gibbs-nipt.cpp:2395-3307 with S = 1, n_gibbs_starts = 1, run_fb_subset = FALSE,
// use_small_eHapsCurrent_tc = FALSE, use_starting_read_labels = TRUE,
// haploid_gibbs_equal_weighting = TRUE, calculate_gamma_on_the_fly irrelevant to results,
// do_block_resampling = FALSE, seed_vector = 0 (functions.R:2566-2678)
int quilt_oracle_gibbs(const QuiltGibbsArgs* a, QuiltGibbsOut* o) {
    if (!a || !o || !a->panel) return QUILT_ERR_BAD_ARG;
    State S;
    setup_state(S, a);
    const int K = S.K, nGrids = S.nGrids, nReads = S.nReads;
    const int nSNPsLocal = a->nSNPs;
    const int n_gibbs_burn_in_its = a->n_gibbs_burn_in_its, n_gibbs_sample_its = a->n_gibbs_sample_its;
    const int n_gibbs_full_its = n_gibbs_burn_in_its + n_gibbs_sample_its;
    const bool record_read_set = (a->flags & QUILT_F_RECORD_READ_SET) != 0;
    const bool gibbs_initialize_iteratively = (a->flags & QUILT_F_GIBBS_INITIALIZE_ITERATIVELY) != 0;
    const bool perform_block_gibbs = (a->flags & QUILT_F_PERFORM_BLOCK_GIBBS) != 0;
    const bool do_shard_block_gibbs = (a->flags & QUILT_F_DO_SHARD_BLOCK_GIBBS) != 0;
    const bool shard_check_every_pair = (a->flags & QUILT_F_SHARD_CHECK_EVERY_PAIR) != 0;
    const bool rare_common = (a->flags & QUILT_F_MAKE_EMATREAD_RARE_COMMON) != 0;
    o->underflow_problem = 0;
    o->underflow_iteration = -1;
    o->n_unif_consumed = 0;
    // episode stream (include/quilt_b200.h): R's unif_rand() stream after the call's first two draws, consumed in the
    // reference's order; pos counts what the reference would have drawn
    const double* stream = a->unif_stream;
    int64_t pos = 0;

    Mat genProbsM_t(3, nSNPsLocal), genProbsF_t(3, nSNPsLocal), hapProbs_t(3, nSNPsLocal);
    Mat genProbsM_t_local(3, nSNPsLocal), genProbsF_t_local(3, nSNPsLocal), hapProbs_t_local(3, nSNPsLocal);
    for (int h = 0; h < 3; h++) {
        // diploid production passes 1x1 buffers for hap 3 (quilt.R:738-745); they are never addressed
        int kk = (h < S.nHaps) ? K : 1, gg = (h < S.nHaps) ? nGrids : 1;
        S.alphaHat_t[h] = Mat(kk, gg);
        S.betaHat_t[h] = Mat(kk, gg);
        S.eMatGrid_t[h] = Mat(kk, gg);
        S.c[h] = Vec(nGrids, 0.0);
    }
    S.H_class.assign(record_read_set ? nReads : 1, 0);
    int n_per_it_likelihoods = n_gibbs_full_its;
    if (n_gibbs_sample_its == 0) n_per_it_likelihoods = 1;
    Mat per_it_likelihoods(n_per_it_likelihoods, 13);
    int i_per_it_likelihoods = 0;

    std::vector<std::vector<int>> rare_per_snp;
    build_eMatRead(S, a, rare_per_snp);
    S.H.assign(a->H0, a->H0 + nReads);
    gibbs_nipt_initialize(S, gibbs_initialize_iteratively);

    int i_result_it;
    int n_results_done = 0;
    int episode = 0;
    for (int iteration = 0; iteration < n_gibbs_full_its; iteration++) {
        if ((iteration + 1) > n_gibbs_burn_in_its) {
            i_result_it = iteration - n_gibbs_burn_in_its;
        } else {
            i_result_it = -2;
        }
        gibbs_nipt_iterate(S, iteration, a->runif_reads, record_read_set, a->class_sum_cutoff, gibbs_initialize_iteratively,
                           a->first_read_for_gibbs_initialization, per_it_likelihoods, i_per_it_likelihoods, i_result_it);
        // underflow check, gibbs-nipt.cpp:2959-2969 (c3 only looked at when ff == 0, sic)
        bool check = true;
        for (int q = 0; q <= 2; q++) {
            if (q == 0) check = is_finite_d(accu2(S.c[0].data(), nGrids));
            if (q == 1) check = is_finite_d(accu2(S.c[1].data(), nGrids));
            if ((a->ff == 0) & (q == 2)) check = is_finite_d(accu2(S.c[2].data(), nGrids));
            if (!check) {
                o->underflow_problem = 1;
                o->underflow_iteration = iteration;
                o->n_unif_consumed = pos;
                return QUILT_OK;
            }
        }
        bool to_block_gibbs = false;
        if (perform_block_gibbs) {
            for (int i = 0; i < a->n_block_gibbs_iterations; i++)
                if (iteration == a->block_gibbs_iterations[i]) to_block_gibbs = true;
        }
        if (perform_block_gibbs & to_block_gibbs) {
            IVec blocked_snps = define_blocked_snps_using_gamma_on_the_fly(
                S, nSNPsLocal, a->smooth_cm, a->shuffle_bin_radius, a->L_grid, a->block_gibbs_quantile_prob,
                (a->flags & QUILT_F_USE_SMOOTH_CM_IN_BLOCK_GIBBS) != 0);
            const double* rb;
            const double* rh;
            if (stream) {
                // six runif(nReads) rows (runif_proposed), runif_block, runif_total (gibbs-nipt.cpp:3013-3018)
                if (pos + 8 * (int64_t)nReads + (S.ff > 0 ? nReads : 0) > a->n_unif_stream) return QUILT_ERR_BAD_ARG;
                rb = stream + pos + 6 * (int64_t)nReads;
                pos += 8 * (int64_t)nReads;
                rh = stream + pos;
            } else {
                rb = a->runif_block + (size_t)episode * nReads;
                rh = a->runif_H_class ? a->runif_H_class + (size_t)episode * nReads : nullptr;
            }
            if (S.ff > 0 && !rh) return QUILT_ERR_BAD_ARG;
            pos += block_gibbs_resampler(S, blocked_snps, rb, rh);
            if (do_shard_block_gibbs) {
                const double* rs;
                if (stream) {
                    if (!shard_check_every_pair) return QUILT_ERR_UNSUPPORTED;  // runif(n_blocks - 1) with data-dependent n_blocks
                    if (pos + (nGrids - 1) > a->n_unif_stream) return QUILT_ERR_BAD_ARG;
                    rs = stream + pos;
                    pos += nGrids - 1;  // block.cpp:2054 with n_blocks = nGrids (:2038-2040)
                } else {
                    rs = a->runif_shard + (size_t)episode * (nGrids - 1);
                }
                int rc = shard_block_gibbs_resampler(S, blocked_snps, shard_check_every_pair, rs);
                if (rc != QUILT_OK) return rc;
            }
            episode++;
        }
        if ((iteration + 1) > n_gibbs_burn_in_its) {
            // unpack_gammas, gibbs-nipt.cpp:2130-2388 (genProbs/hapProbs branch + equal-weight fly weighter :2063-2091)
            if (rare_common) {
                calculate_genProbs_and_hapProbs_final_rare_common(S, a->panel, a->which_haps_to_use, genProbsM_t_local,
                                                                  genProbsF_t_local, hapProbs_t_local, rare_per_snp);
            } else {
                calculate_genProbs_and_hapProbs_using_binary_objects(S, a->panel, a->which_haps_to_use, genProbsM_t_local,
                                                                     genProbsF_t_local, hapProbs_t_local);
            }
            if (i_result_it + 1 == 1) {
                genProbsM_t = genProbsM_t_local;
                genProbsF_t = genProbsF_t_local;
                hapProbs_t = hapProbs_t_local;
            } else {
                double relative_difference = std::exp(1.0 - 1.0);
                for (size_t i = 0; i < genProbsM_t.d.size(); i++) {
                    genProbsM_t.d[i] += relative_difference * genProbsM_t_local.d[i];
                    genProbsF_t.d[i] += relative_difference * genProbsF_t_local.d[i];
                    hapProbs_t.d[i] += relative_difference * hapProbs_t_local.d[i];
                }
            }
            // list_of_ending_read_labels.push_back(clone(H), "H") (gibbs-nipt.cpp:3104)
            if (o->H_sample_its)
                for (int r = 0; r < nReads; r++) o->H_sample_its[(size_t)i_result_it * nReads + r] = S.H[r];
            n_results_done++;
        }
    }
    // gibbs-nipt.cpp:3145-3197
    double gamma_temp;
    int n_results = n_gibbs_sample_its;
    if (n_results > 0) {
        gamma_temp = 1 / double(n_results);
    } else {
        gamma_temp = 1;
    }
    if (gamma_temp != 1) {
        for (size_t i = 0; i < genProbsM_t.d.size(); i++) {
            genProbsM_t.d[i] *= gamma_temp;
            genProbsF_t.d[i] *= gamma_temp;
            hapProbs_t.d[i] *= gamma_temp;
        }
    }
    if (o->hapProbs_t) std::memcpy(o->hapProbs_t, hapProbs_t.d.data(), sizeof(double) * hapProbs_t.d.size());
    if (o->genProbsM_t) std::memcpy(o->genProbsM_t, genProbsM_t.d.data(), sizeof(double) * genProbsM_t.d.size());
    if (o->genProbsF_t) std::memcpy(o->genProbsF_t, genProbsF_t.d.data(), sizeof(double) * genProbsF_t.d.size());
    if (o->H)
        for (int r = 0; r < nReads; r++) o->H[r] = S.H[r];
    if (o->H_class && record_read_set)
        for (int r = 0; r < nReads; r++) o->H_class[r] = S.H_class[r];
    if (o->per_it_likelihoods)
        std::memcpy(o->per_it_likelihoods, per_it_likelihoods.d.data(), sizeof(double) * per_it_likelihoods.d.size());
    if (a->flags & QUILT_F_RETURN_ALPHA) {
        for (int h = 0; h < S.nHaps; h++) {
            if (o->alphaHat_t[h]) std::memcpy(o->alphaHat_t[h], S.alphaHat_t[h].d.data(), sizeof(double) * (size_t)K * nGrids);
            if (o->betaHat_t[h]) std::memcpy(o->betaHat_t[h], S.betaHat_t[h].d.data(), sizeof(double) * (size_t)K * nGrids);
            if (o->eMatGrid_t[h]) std::memcpy(o->eMatGrid_t[h], S.eMatGrid_t[h].d.data(), sizeof(double) * (size_t)K * nGrids);
            if (o->c[h]) std::memcpy(o->c[h], S.c[h].data(), sizeof(double) * nGrids);
        }
    }
    if ((a->flags & QUILT_F_RETURN_EXTRA) && o->eMatRead_t)
        std::memcpy(o->eMatRead_t, S.eMatRead_t.d.data(), sizeof(double) * (size_t)K * nReads);
    if (o->read_category)
        for (int r = 0; r < nReads; r++) o->read_category[r] = S.read_category[r];
    (void)n_results_done;
    o->n_unif_consumed = stream ? pos : 0;
    return QUILT_OK;
}

int quilt_oracle_make_eMatRead_t(const QuiltGibbsArgs* a, double* eMatRead_t, int32_t* read_category) {
    if (!a || !a->panel || !eMatRead_t) return QUILT_ERR_BAD_ARG;
    State S;
    setup_state(S, a);
    std::vector<std::vector<int>> rare_per_snp;
    build_eMatRead(S, a, rare_per_snp);
    std::memcpy(eMatRead_t, S.eMatRead_t.d.data(), sizeof(double) * (size_t)S.K * S.nReads);
    if (read_category)
        for (int r = 0; r < S.nReads; r++) read_category[r] = S.read_category[r];
    return QUILT_OK;
}

// the 32-way unpack primitive behind rcpp_int_expand / inflate_fhb (copied-from-stitch.cpp:50-108): here the packed
// words themselves for the selected haplotypes, LSB = first SNP of the grid.  all_snps = 1 assembles words over the
// all-SNP axis from the common-SNP panel plus rare_per_hap_info (rare_common.R:202-322).
int quilt_oracle_unpack_panel(const QuiltPanel* p, int32_t K, const int32_t* which, int32_t all_snps, uint32_t* words) {
    if (!p || !which || !words) return QUILT_ERR_BAD_ARG;
    if (!all_snps) {
        for (int g = 0; g < p->nGrids; g++)
            for (int k = 0; k < K; k++) words[(size_t)g * K + k] = (uint32_t)panel_word(p, which[k] - 1, g);
        return QUILT_OK;
    }
    int nG = (p->nSNPs_all + 31) / 32;
    std::fill(words, words + (size_t)nG * K, 0u);
    for (int s = 0; s < p->nSNPs_all; s++) {
        if (!p->snp_is_common[s]) continue;
        int cs = p->common_snp_index[s] - 1;
        for (int k = 0; k < K; k++) {
            uint32_t w = (uint32_t)panel_word(p, which[k] - 1, cs / 32);
            if ((w >> (cs % 32)) & 1u) words[(size_t)(s / 32) * K + k] |= (1u << (s % 32));
        }
    }
    for (int k = 0; k < K; k++) {
        int h = which[k] - 1;
        for (int64_t j = p->rare_hap_offsets[h]; j < p->rare_hap_offsets[h + 1]; j++) {
            int s = p->rare_hap_snps[j] - 1;
            words[(size_t)(s / 32) * K + k] |= (1u << (s % 32));
        }
    }
    return QUILT_OK;
}

// Rcpp_run_forward_haploid + Rcpp_run_backward_haploid with uniform prior (copied-from-stitch.cpp:340-409,
// wrapper gibbs-nipt.cpp:453-487)
int quilt_oracle_forward_backward(int32_t K, int32_t nGrids, const double* eMatGrid_t, const double* transMatRate_tc_H,
                                  double* alphaHat_t, double* betaHat_t, double* c) {
    Mat e(K, nGrids), a(K, nGrids), b(K, nGrids);
    std::memcpy(e.d.data(), eMatGrid_t, sizeof(double) * (size_t)K * nGrids);
    Vec cc(nGrids, 0.0);
    Trans tm{transMatRate_tc_H};
    const double prior = 1 / double(K);
    run_forward_haploid(a, cc, e, prior, tm, false);
    double* bl = b.col(nGrids - 1);
    for (int k = 0; k < K; k++) bl[k] = cc[nGrids - 1];
    run_backward_haploid(b, cc, e, prior, tm);
    std::memcpy(alphaHat_t, a.d.data(), sizeof(double) * (size_t)K * nGrids);
    std::memcpy(betaHat_t, b.d.data(), sizeof(double) * (size_t)K * nGrids);
    std::memcpy(c, cc.data(), sizeof(double) * nGrids);
    return QUILT_OK;
}

// extra oracle-only hooks used by tests/test_oracle_invariants.py -------------------------------------------------
// generic backward and fast backward on the same inputs (test-unit-gibbs-nipt-parts.R:177-264)
int quilt_oracle_backward_pair(int32_t K, int32_t nGrids, const double* eMatGrid_t, const double* transMatRate_tc_H, const double* c,
                               const uint8_t* grid_has_read, double* beta_generic, double* beta_fast) {
    Mat e(K, nGrids), b1(K, nGrids), b2(K, nGrids);
    std::memcpy(e.d.data(), eMatGrid_t, sizeof(double) * (size_t)K * nGrids);
    Vec cc(c, c + nGrids);
    Trans tm{transMatRate_tc_H};
    std::vector<uint8_t> ghr(grid_has_read, grid_has_read + nGrids);
    for (int k = 0; k < K; k++) b1(k, nGrids - 1) = b2(k, nGrids - 1) = cc[nGrids - 1];
    run_backward_haploid(b1, cc, e, 1 / double(K), tm);
    run_backward_haploid_QUILT_faster(b2, cc, e, tm, ghr);
    std::memcpy(beta_generic, b1.d.data(), sizeof(double) * (size_t)K * nGrids);
    std::memcpy(beta_fast, b2.d.data(), sizeof(double) * (size_t)K * nGrids);
    return QUILT_OK;
}

// incremental forward (both flavours) from a full forward's column g-1 (test-unit-gibbs-nipt-parts.R:1-174)
int quilt_oracle_forward_one_pair(int32_t K, int32_t nGrids, const double* eMatGrid_t, const double* transMatRate_tc_H,
                                  const uint8_t* grid_has_read, int32_t iGrid, double* alphaHat_t /*in/out*/, double* c /*in/out*/,
                                  int32_t faster) {
    Mat e(K, nGrids), a(K, nGrids);
    std::memcpy(e.d.data(), eMatGrid_t, sizeof(double) * (size_t)K * nGrids);
    std::memcpy(a.d.data(), alphaHat_t, sizeof(double) * (size_t)K * nGrids);
    Vec cc(c, c + nGrids);
    Trans tm{transMatRate_tc_H};
    std::vector<uint8_t> ghr(grid_has_read, grid_has_read + nGrids);
    double mlc = 0;
    if (faster)
        alpha_forward_one_QUILT_faster(iGrid, K, a, tm, e, cc, mlc, ghr, true);
    else
        alpha_forward_one(iGrid, K, a, tm, e, 1 / double(K), cc, mlc, true);
    std::memcpy(alphaHat_t, a.d.data(), sizeof(double) * (size_t)K * nGrids);
    std::memcpy(c, cc.data(), sizeof(double) * nGrids);
    return QUILT_OK;
}

}  // extern "C"
