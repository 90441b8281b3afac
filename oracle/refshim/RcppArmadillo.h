// RcppArmadillo.h — header-only STAND-IN for <RcppArmadillo.h> (TEST INFRASTRUCTURE, not product code).
//
// Purpose: compile the UNMODIFIED reference sources under /root/reference/QUILT/src (copied-from-stitch.cpp,
// gibbs-small.cpp, gibbs-nipt.cpp, gibbs-nipt-block.cpp, reference-single.cpp) in an image that has neither R
// nor Rcpp nor Armadillo, so that the reference's own code — not a restatement — produces the vectors the
// oracle and the CUDA path are checked against (oracle/_ref/libquiltref.so, recipe: oracle/refshim/Makefile).
//
// Nothing here is copied from Rcpp or Armadillo; it re-implements the small API subset those five files use,
// with the numerical conventions that matter for fp64 parity kept identical to the real libraries:
//   * arma containers zero-fill on construction (Armadillo >= 10.5 behaviour, which the reference relies on);
//   * element-wise expressions are evaluated lazily, one element at a time, each binary op rounded separately;
//   * accu()/sum() of a vector (expression) use two interleaved accumulators (even / odd elements) added at
//     the end — Armadillo's arrayops::accumulate / accu_proxy_linear in non-fast-math builds;
//   * sort_index() is std::sort over {value, index} packets with a strict comparator (not stable), as in
//     Armadillo's arma_sort_index_helper;
//   * Rcpp sugar sum() is a plain left-to-right loop; Rcpp::max returns on the first NaN met;
//   * Rcpp::runif / Rcpp::sample consume a unif_rand() stream exactly as Rcpp's sugar does (runif rejects
//     values outside (0,1); sample(n, 1) = int(n * u + 1); sample(x, 1, false, probs) = Normalize + revsort +
//     cumulative scan) — the stream itself is injected through refshim::rng().
#ifndef REFSHIM_RCPPARMADILLO_H
#define REFSHIM_RCPPARMADILLO_H

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <initializer_list>
#include <iostream>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

// =====================================================================================================
//                                               arma
// =====================================================================================================
namespace Rcpp { template <int RTYPE> class Vector; template <int RTYPE> class Matrix; }

namespace arma {

typedef unsigned long long uword;
typedef long long sword;

namespace fill {
struct zeros_t {};
struct ones_t {};
struct none_t {};
static const zeros_t zeros = zeros_t();
static const ones_t ones = ones_t();
static const none_t none = none_t();
}  // namespace fill

template <class eT> class Mat;
template <class eT> class Col;
template <class eT> class Row;
template <class eT> class Cube;
template <class eT> class subview_col;
template <class eT> class subview_row;

// CRTP base of everything that can be read element by element (column-major linear index)
struct BaseTag {};
template <class D>
struct Base : public BaseTag {
    const D& get() const { return static_cast<const D&>(*this); }
};
template <class T> struct is_expr { static const bool value = std::is_base_of<BaseTag, T>::value; };
#define REFSHIM_IF_EXPR(T, ...) typename std::enable_if<is_expr<T>::value, __VA_ARGS__>::type
#define REFSHIM_IF_EXPR2(A, B, ...) typename std::enable_if<is_expr<A>::value && is_expr<B>::value, __VA_ARGS__>::type

// how an operand is held inside an expression node: containers by reference, light views by value
template <class T> struct hold { typedef T type; };
template <class eT> struct hold<Mat<eT> > { typedef const Mat<eT>& type; };
template <class eT> struct hold<Col<eT> > { typedef const Col<eT>& type; };
template <class eT> struct hold<Row<eT> > { typedef const Row<eT>& type; };

// compile-time shape class: 0 = matrix, 1 = column vector, 2 = row vector
template <class T> struct shape_of { static const int value = T::shape; };

struct op_plus  { template <class T> static T apply(T a, T b) { return a + b; } };
struct op_minus { template <class T> static T apply(T a, T b) { return a - b; } };
struct op_schur { template <class T> static T apply(T a, T b) { return a * b; } };
struct op_div   { template <class T> static T apply(T a, T b) { return a / b; } };
// scalar forms: "post" = expr (op) k, "pre" = k (op) expr
struct sop_plus      { template <class T> static T apply(T a, T k) { return a + k; } };
struct sop_minus_post{ template <class T> static T apply(T a, T k) { return a - k; } };
struct sop_minus_pre { template <class T> static T apply(T a, T k) { return k - a; } };
struct sop_times     { template <class T> static T apply(T a, T k) { return a * k; } };
struct sop_div_post  { template <class T> static T apply(T a, T k) { return a / k; } };
struct sop_div_pre   { template <class T> static T apply(T a, T k) { return k / a; } };
struct sop_neg       { template <class T> static T apply(T a, T)   { return -a; } };
struct sop_log       { template <class T> static T apply(T a, T)   { return std::log(a); } };
struct sop_exp       { template <class T> static T apply(T a, T)   { return std::exp(a); } };
struct sop_sqrt      { template <class T> static T apply(T a, T)   { return std::sqrt(a); } };
struct sop_abs       { template <class T> static T apply(T a, T)   { return std::abs(a); } };

template <class A, class B, class Op>
struct eGlue : public Base<eGlue<A, B, Op> > {
    typedef typename A::elem_type elem_type;
    static const int shape = A::shape;
    typename hold<A>::type a;
    typename hold<B>::type b;
    eGlue(const A& a_, const B& b_) : a(a_), b(b_) {
        if (a.get_n_elem() != b.get_n_elem()) throw std::logic_error("refshim arma: element-wise op on objects of different size");
    }
    uword get_n_rows() const { return a.get_n_rows(); }
    uword get_n_cols() const { return a.get_n_cols(); }
    uword get_n_elem() const { return a.get_n_elem(); }
    elem_type operator[](uword i) const { return Op::apply(a[i], b[i]); }
};

template <class A, class Op>
struct eOp : public Base<eOp<A, Op> > {
    typedef typename A::elem_type elem_type;
    static const int shape = A::shape;
    typename hold<A>::type a;
    elem_type k;
    eOp(const A& a_, elem_type k_) : a(a_), k(k_) {}
    uword get_n_rows() const { return a.get_n_rows(); }
    uword get_n_cols() const { return a.get_n_cols(); }
    uword get_n_elem() const { return a.get_n_elem(); }
    elem_type operator[](uword i) const { return Op::apply(a[i], k); }
};

// generator expressions: zeros(r, c), ones(r, c)
template <class eT>
struct Gen : public Base<Gen<eT> > {
    typedef eT elem_type;
    static const int shape = 0;
    uword r, c;
    eT v;
    Gen(uword r_, uword c_, eT v_) : r(r_), c(c_), v(v_) {}
    uword get_n_rows() const { return r; }
    uword get_n_cols() const { return c; }
    uword get_n_elem() const { return r * c; }
    eT operator[](uword) const { return v; }
};
template <class eT>
struct GenCube {
    uword r, c, s;
    eT v;
};

// ---- operators on expressions
#define REFSHIM_BINOP(sym, OP)                                                                           \
    template <class A, class B>                                                                            \
    inline REFSHIM_IF_EXPR2(A, B, eGlue<A, B, OP>) operator sym(const A& a, const B& b) { return eGlue<A, B, OP>(a, b); }
REFSHIM_BINOP(+, op_plus)
REFSHIM_BINOP(-, op_minus)
REFSHIM_BINOP(%, op_schur)
REFSHIM_BINOP(/, op_div)
#undef REFSHIM_BINOP

template <class T> struct is_scalar { static const bool value = std::is_arithmetic<T>::value; };

#define REFSHIM_SCALAR_POST(sym, OP)                                                                       \
    template <class A, class S>                                                                            \
    inline typename std::enable_if<is_scalar<S>::value && is_expr<A>::value, eOp<A, OP> >::type operator sym(const A& a, S k) { \
        return eOp<A, OP>(a, (typename A::elem_type)k);                                              \
    }
#define REFSHIM_SCALAR_PRE(sym, OP)                                                                        \
    template <class A, class S>                                                                            \
    inline typename std::enable_if<is_scalar<S>::value && is_expr<A>::value, eOp<A, OP> >::type operator sym(S k, const A& a) { \
        return eOp<A, OP>(a, (typename A::elem_type)k);                                              \
    }
REFSHIM_SCALAR_POST(+, sop_plus)
REFSHIM_SCALAR_PRE(+, sop_plus)
REFSHIM_SCALAR_POST(-, sop_minus_post)
REFSHIM_SCALAR_PRE(-, sop_minus_pre)
REFSHIM_SCALAR_POST(*, sop_times)
REFSHIM_SCALAR_PRE(*, sop_times)
REFSHIM_SCALAR_POST(/, sop_div_post)
REFSHIM_SCALAR_PRE(/, sop_div_pre)
#undef REFSHIM_SCALAR_POST
#undef REFSHIM_SCALAR_PRE

template <class A> inline REFSHIM_IF_EXPR(A, eOp<A, sop_neg>) operator-(const A& a) { return eOp<A, sop_neg>(a, 0); }
template <class A> inline REFSHIM_IF_EXPR(A, eOp<A, sop_log>) log(const A& a) { return eOp<A, sop_log>(a, 0); }
template <class A> inline REFSHIM_IF_EXPR(A, eOp<A, sop_exp>) exp(const A& a) { return eOp<A, sop_exp>(a, 0); }
template <class A> inline REFSHIM_IF_EXPR(A, eOp<A, sop_sqrt>) sqrt(const A& a) { return eOp<A, sop_sqrt>(a, 0); }
template <class A> inline REFSHIM_IF_EXPR(A, eOp<A, sop_abs>) abs(const A& a) { return eOp<A, sop_abs>(a, 0); }

// ---- in-place helpers shared by Mat and the views (D needs operator[] returning a reference-like lvalue via ref(i))
template <class Dst, class Src>
inline void assign_from(Dst& d, const Src& s, const char* what) {
    const uword n = d.get_n_elem();
    if (s.get_n_elem() != n) throw std::logic_error(std::string("refshim arma: size mismatch in ") + what);
    for (uword i = 0; i < n; ++i) d.ref(i) = s[i];
}

// ------------------------------------------------------------------------------------------ subview_col
template <class eT>
class subview_col : public Base<subview_col<eT> > {
public:
    typedef eT elem_type;
    static const int shape = 1;
    eT* colmem;
    uword n_rows;
    static const uword n_cols = 1;
    uword n_elem;
    subview_col(eT* p, uword n) : colmem(p), n_rows(n), n_elem(n) {}
    subview_col(const subview_col& o) : Base<subview_col<eT> >(), colmem(o.colmem), n_rows(o.n_rows), n_elem(o.n_elem) {}
    uword get_n_rows() const { return n_rows; }
    uword get_n_cols() const { return 1; }
    uword get_n_elem() const { return n_elem; }
    eT operator[](uword i) const { return colmem[i]; }
    eT& ref(uword i) { return colmem[i]; }
    eT& operator()(uword i) { return colmem[i]; }
    eT operator()(uword i) const { return colmem[i]; }
    eT& at(uword i) { return colmem[i]; }
    uword size() const { return n_elem; }
    void fill(eT v) { for (uword i = 0; i < n_elem; ++i) colmem[i] = v; }
    void zeros() { fill(eT(0)); }
    void ones() { fill(eT(1)); }
    // assignment copies VALUES into the viewed memory
    subview_col& operator=(const subview_col& o) {
        if (o.n_elem != n_elem) throw std::logic_error("refshim arma: size mismatch in subview_col=");
        if (colmem != o.colmem) std::memmove(colmem, o.colmem, sizeof(eT) * n_elem);
        return *this;
    }
    template <class E> subview_col& operator=(const Base<E>& e) { assign_from(*this, e.get(), "subview_col="); return *this; }
    subview_col& operator=(eT v) { if (n_elem != 1) throw std::logic_error("refshim arma: scalar = on subview"); colmem[0] = v; return *this; }
#define REFSHIM_INPLACE(sym)                                                                               \
    template <class E> subview_col& operator sym(const Base<E>& e) {                                       \
        const E& s = e.get();                                                                              \
        if (s.get_n_elem() != n_elem) throw std::logic_error("refshim arma: size mismatch in subview_col in-place op"); \
        for (uword i = 0; i < n_elem; ++i) colmem[i] sym s[i];                                             \
        return *this;                                                                                      \
    }                                                                                                      \
    subview_col& operator sym(eT k) { for (uword i = 0; i < n_elem; ++i) colmem[i] sym k; return *this; }
    REFSHIM_INPLACE(+=)
    REFSHIM_INPLACE(-=)
    REFSHIM_INPLACE(*=)
    REFSHIM_INPLACE(/=)
#undef REFSHIM_INPLACE
    template <class E> subview_col& operator%=(const Base<E>& e) { return (*this) *= e; }
    bool is_finite() const { for (uword i = 0; i < n_elem; ++i) if (!std::isfinite((double)colmem[i])) return false; return true; }
    eT* begin() { return colmem; }
    eT* end() { return colmem + n_elem; }
    const eT* begin() const { return colmem; }
    const eT* end() const { return colmem + n_elem; }
};

// ------------------------------------------------------------------------------------------ subview_row (strided)
template <class eT>
class subview_row : public Base<subview_row<eT> > {
public:
    typedef eT elem_type;
    static const int shape = 2;
    eT* mem0;
    uword stride;
    static const uword n_rows = 1;
    uword n_cols;
    uword n_elem;
    subview_row(eT* p, uword stride_, uword n) : mem0(p), stride(stride_), n_cols(n), n_elem(n) {}
    subview_row(const subview_row& o) : Base<subview_row<eT> >(), mem0(o.mem0), stride(o.stride), n_cols(o.n_cols), n_elem(o.n_elem) {}
    uword get_n_rows() const { return 1; }
    uword get_n_cols() const { return n_cols; }
    uword get_n_elem() const { return n_elem; }
    eT operator[](uword i) const { return mem0[i * stride]; }
    eT& ref(uword i) { return mem0[i * stride]; }
    eT& operator()(uword i) { return mem0[i * stride]; }
    eT operator()(uword i) const { return mem0[i * stride]; }
    void fill(eT v) { for (uword i = 0; i < n_elem; ++i) ref(i) = v; }
    subview_row& operator=(const subview_row& o) {
        if (o.n_elem != n_elem) throw std::logic_error("refshim arma: size mismatch in subview_row=");
        std::vector<eT> tmp(n_elem);
        for (uword i = 0; i < n_elem; ++i) tmp[i] = o[i];
        for (uword i = 0; i < n_elem; ++i) ref(i) = tmp[i];
        return *this;
    }
    template <class E> subview_row& operator=(const Base<E>& e) { assign_from(*this, e.get(), "subview_row="); return *this; }
#define REFSHIM_INPLACE(sym)                                                                               \
    template <class E> subview_row& operator sym(const Base<E>& e) {                                       \
        const E& s = e.get();                                                                              \
        if (s.get_n_elem() != n_elem) throw std::logic_error("refshim arma: size mismatch in subview_row in-place op"); \
        for (uword i = 0; i < n_elem; ++i) ref(i) sym s[i];                                                \
        return *this;                                                                                      \
    }                                                                                                      \
    subview_row& operator sym(eT k) { for (uword i = 0; i < n_elem; ++i) ref(i) sym k; return *this; }
    REFSHIM_INPLACE(+=)
    REFSHIM_INPLACE(-=)
    REFSHIM_INPLACE(*=)
    REFSHIM_INPLACE(/=)
#undef REFSHIM_INPLACE
    template <class E> subview_row& operator%=(const Base<E>& e) { return (*this) *= e; }
};

// ------------------------------------------------------------------------------------------ Mat
struct SizeMat { uword n_rows, n_cols; };

template <class eT>
class Mat : public Base<Mat<eT> > {
public:
    typedef eT elem_type;
    static const int shape = 0;
    uword n_rows, n_cols, n_elem;
    eT* mem;

protected:
    std::vector<eT> own;   // owned storage (empty when the matrix wraps foreign memory)
    bool foreign;
    bool strict_foreign = false;
    int vec_state;         // 0 matrix, 1 column vector, 2 row vector (fixes the orientation on resize / assignment)

    void init_owned(uword r, uword c, eT v) {
        foreign = false;
        n_rows = r; n_cols = c; n_elem = r * c;
        own.assign((size_t)n_elem, v);
        mem = own.data();
    }

public:
    Mat() : n_rows(0), n_cols(0), n_elem(0), mem(nullptr), foreign(false), vec_state(0) {}
    Mat(uword r, uword c) : vec_state(0) { init_owned(r, c, eT(0)); }
    Mat(uword r, uword c, fill::zeros_t) : vec_state(0) { init_owned(r, c, eT(0)); }
    Mat(uword r, uword c, fill::ones_t) : vec_state(0) { init_owned(r, c, eT(1)); }
    Mat(uword r, uword c, fill::none_t) : vec_state(0) { init_owned(r, c, eT(0)); }
    // wrap foreign memory (Armadillo's advanced constructor with copy_aux_mem = false)
    Mat(eT* aux, uword r, uword c, bool copy_aux_mem = true, bool strict = false) : vec_state(0) {
        strict_foreign = strict && !copy_aux_mem;
        if (copy_aux_mem) {
            init_owned(r, c, eT(0));
            if (n_elem) std::memcpy(mem, aux, sizeof(eT) * n_elem);
        } else {
            foreign = true; n_rows = r; n_cols = c; n_elem = r * c; mem = aux;
        }
    }
    Mat(const Mat& o) : Base<Mat<eT> >(), vec_state(0) {
        init_owned(o.n_rows, o.n_cols, eT(0));
        if (n_elem) std::memcpy(mem, o.mem, sizeof(eT) * n_elem);
    }
    Mat(Mat&& o) : Base<Mat<eT> >(), vec_state(0) {
        if (o.foreign) { foreign = true; n_rows = o.n_rows; n_cols = o.n_cols; n_elem = o.n_elem; mem = o.mem; }
        else { foreign = false; own.swap(o.own); n_rows = o.n_rows; n_cols = o.n_cols; n_elem = o.n_elem; mem = own.data(); }
        o.n_rows = o.n_cols = o.n_elem = 0; o.mem = nullptr; o.own.clear();
    }
    template <int RT> Mat(const Rcpp::Vector<RT>& v);   // RcppArmadillo extension: column vector from an R vector
    template <int RT> Mat(const Rcpp::Matrix<RT>& v);
    template <class E> Mat(const Base<E>& e) : vec_state(0) {
        const E& s = e.get();
        init_owned(s.get_n_rows(), s.get_n_cols(), eT(0));
        for (uword i = 0; i < n_elem; ++i) mem[i] = s[i];
    }
    virtual ~Mat() {}

    uword get_n_rows() const { return n_rows; }
    uword get_n_cols() const { return n_cols; }
    uword get_n_elem() const { return n_elem; }
    eT operator[](uword i) const { return mem[i]; }
    eT& operator[](uword i) { return mem[i]; }
    eT& ref(uword i) { return mem[i]; }
    eT& operator()(uword i) { return mem[i]; }
    eT operator()(uword i) const { return mem[i]; }
    eT& operator()(uword i, uword j) { return mem[i + j * n_rows]; }
    eT operator()(uword i, uword j) const { return mem[i + j * n_rows]; }
    eT& at(uword i) { return mem[i]; }
    eT at(uword i) const { return mem[i]; }
    eT& at(uword i, uword j) { return mem[i + j * n_rows]; }
    eT at(uword i, uword j) const { return mem[i + j * n_rows]; }
    uword size() const { return n_elem; }
    eT* memptr() { return mem; }
    const eT* memptr() const { return mem; }
    eT* begin() { return mem; }
    eT* end() { return mem + n_elem; }
    const eT* begin() const { return mem; }
    const eT* end() const { return mem + n_elem; }
    bool is_empty() const { return n_elem == 0; }

    // resize keeping nothing (Armadillo: set_size leaves memory uninitialised; zero here is a superset)
    void set_size(uword r, uword c) {
        if (r == n_rows && c == n_cols) return;
        if (foreign && r * c == n_elem) { n_rows = r; n_cols = c; return; }
        if (foreign && strict_foreign) throw std::logic_error("refshim arma: cannot resize a matrix that wraps foreign memory (strict)");
        init_owned(r, c, eT(0));   // a non-strict foreign matrix lets go of the foreign memory, like Armadillo's
    }
    void set_size(uword n) { if (vec_state == 2) set_size(1, n); else set_size(n, 1); }
    void zeros() { fill(eT(0)); }
    void zeros(uword r, uword c) { set_size(r, c); fill(eT(0)); }
    void ones() { fill(eT(1)); }
    void ones(uword r, uword c) { set_size(r, c); fill(eT(1)); }
    Mat& fill(eT v) { for (uword i = 0; i < n_elem; ++i) mem[i] = v; return *this; }

    subview_col<eT> col(uword j) { return subview_col<eT>(mem + j * n_rows, n_rows); }
    const subview_col<eT> col(uword j) const { return subview_col<eT>(const_cast<eT*>(mem) + j * n_rows, n_rows); }
    subview_row<eT> row(uword i) { return subview_row<eT>(mem + i, n_rows, n_cols); }
    const subview_row<eT> row(uword i) const { return subview_row<eT>(const_cast<eT*>(mem) + i, n_rows, n_cols); }

    // assignment: into foreign memory the shape must match (R-owned buffers are written in place)
    template <class E> void assign_expr(const E& s) {
        uword r = s.get_n_rows(), c = s.get_n_cols();
        if (vec_state == 1 && r == 1 && c != 1) { r = c; c = 1; }
        if (vec_state == 2 && c == 1 && r != 1) { c = r; r = 1; }
        set_size(r, c);
        for (uword i = 0; i < n_elem; ++i) mem[i] = s[i];
    }
    Mat& operator=(const Mat& o) {
        if (this == &o) return *this;
        if (o.n_elem && mem && o.mem >= mem && o.mem < mem + n_elem && n_elem != o.n_elem) throw std::logic_error("refshim arma: aliasing assignment");
        assign_expr(o);
        return *this;
    }
    Mat& operator=(Mat&& o) {
        if (this == &o) return *this;
        if (!foreign && !o.foreign && vec_state == 0) {
            own.swap(o.own); n_rows = o.n_rows; n_cols = o.n_cols; n_elem = o.n_elem; mem = own.data();
            o.n_rows = o.n_cols = o.n_elem = 0; o.mem = nullptr; o.own.clear();
        } else {
            assign_expr(o);
        }
        return *this;
    }
    template <class E> Mat& operator=(const Base<E>& e) {
        // evaluate into a temporary first when shapes differ (possible alias), else in place (element-wise safe)
        const E& s = e.get();
        uword r = s.get_n_rows(), c = s.get_n_cols();
        if (r * c == n_elem && n_elem > 0 && ((r == n_rows && c == n_cols) || vec_state != 0)) {
            for (uword i = 0; i < n_elem; ++i) mem[i] = s[i];
        } else {
            std::vector<eT> tmp((size_t)(r * c));
            for (uword i = 0; i < r * c; ++i) tmp[i] = s[i];
            if (vec_state == 1 && r == 1 && c != 1) { r = c; c = 1; }
            if (vec_state == 2 && c == 1 && r != 1) { c = r; r = 1; }
            set_size(r, c);
            for (uword i = 0; i < n_elem; ++i) mem[i] = tmp[i];
        }
        return *this;
    }
#define REFSHIM_INPLACE(sym)                                                                               \
    template <class E> Mat& operator sym(const Base<E>& e) {                                               \
        const E& s = e.get();                                                                              \
        if (s.get_n_elem() != n_elem) throw std::logic_error("refshim arma: size mismatch in Mat in-place op"); \
        for (uword i = 0; i < n_elem; ++i) mem[i] sym s[i];                                                \
        return *this;                                                                                      \
    }                                                                                                      \
    Mat& operator sym(eT k) { for (uword i = 0; i < n_elem; ++i) mem[i] sym k; return *this; }
    REFSHIM_INPLACE(+=)
    REFSHIM_INPLACE(-=)
    REFSHIM_INPLACE(/=)
#undef REFSHIM_INPLACE
    Mat& operator*=(eT k) { for (uword i = 0; i < n_elem; ++i) mem[i] *= k; return *this; }
    template <class E> Mat& operator%=(const Base<E>& e) {
        const E& s = e.get();
        if (s.get_n_elem() != n_elem) throw std::logic_error("refshim arma: size mismatch in Mat %=");
        for (uword i = 0; i < n_elem; ++i) mem[i] *= s[i];
        return *this;
    }
    bool is_finite() const { for (uword i = 0; i < n_elem; ++i) if (!std::isfinite((double)mem[i])) return false; return true; }
    bool has_nan() const { for (uword i = 0; i < n_elem; ++i) if (std::isnan((double)mem[i])) return true; return false; }
    eT max() const { if (!n_elem) throw std::logic_error("max(): object has no elements"); eT m = mem[0]; for (uword i = 1; i < n_elem; ++i) if (mem[i] > m) m = mem[i]; return m; }
    eT min() const { if (!n_elem) throw std::logic_error("min(): object has no elements"); eT m = mem[0]; for (uword i = 1; i < n_elem; ++i) if (mem[i] < m) m = mem[i]; return m; }
};

template <class eT>
class Col : public Mat<eT> {
public:
    static const int shape = 1;
    typedef eT elem_type;
    Col() : Mat<eT>() { this->vec_state = 1; this->n_cols = 1; }
    explicit Col(uword n) : Mat<eT>(n, 1) { this->vec_state = 1; }
    Col(uword n, fill::zeros_t) : Mat<eT>(n, 1) { this->vec_state = 1; }
    Col(uword n, fill::ones_t f) : Mat<eT>(n, 1, f) { this->vec_state = 1; }
    Col(uword r, uword c) : Mat<eT>(r, c) { this->vec_state = 1; }
    Col(eT* aux, uword n, bool copy_aux_mem = true, bool strict = false) : Mat<eT>(aux, n, 1, copy_aux_mem, strict) { this->vec_state = 1; }
    Col(const Col& o) : Mat<eT>(static_cast<const Mat<eT>&>(o)) { this->vec_state = 1; }
    Col(Col&& o) : Mat<eT>(static_cast<Mat<eT>&&>(o)) { this->vec_state = 1; }
    template <int RT> Col(const Rcpp::Vector<RT>& v);
    template <class E> Col(const Base<E>& e) : Mat<eT>() { this->vec_state = 1; this->n_cols = 1; Mat<eT>::operator=(e); }
    Col& operator=(const Col& o) { Mat<eT>::operator=(static_cast<const Mat<eT>&>(o)); return *this; }
    Col& operator=(Col&& o) { Mat<eT>::operator=(static_cast<const Mat<eT>&>(o)); return *this; }
    template <class E> Col& operator=(const Base<E>& e) { Mat<eT>::operator=(e); return *this; }
    subview_col<eT> subvec(uword a, uword b) { return subview_col<eT>(this->mem + a, b - a + 1); }
    subview_col<eT> head(uword n) { return subview_col<eT>(this->mem, n); }
    subview_col<eT> tail(uword n) { return subview_col<eT>(this->mem + (this->n_elem - n), n); }
};

template <class eT>
class Row : public Mat<eT> {
public:
    static const int shape = 2;
    typedef eT elem_type;
    Row() : Mat<eT>() { this->vec_state = 2; this->n_rows = 1; }
    explicit Row(uword n) : Mat<eT>(1, n) { this->vec_state = 2; }
    Row(uword n, fill::zeros_t) : Mat<eT>(1, n) { this->vec_state = 2; }
    Row(uword n, fill::ones_t f) : Mat<eT>(1, n, f) { this->vec_state = 2; }
    Row(uword r, uword c) : Mat<eT>(r, c) { this->vec_state = 2; }
    Row(eT* aux, uword n, bool copy_aux_mem = true, bool strict = false) : Mat<eT>(aux, 1, n, copy_aux_mem, strict) { this->vec_state = 2; }
    Row(const Row& o) : Mat<eT>(static_cast<const Mat<eT>&>(o)) { this->vec_state = 2; }
    Row(Row&& o) : Mat<eT>(static_cast<Mat<eT>&&>(o)) { this->vec_state = 2; }
    template <int RT> Row(const Rcpp::Vector<RT>& v);
    template <class E> Row(const Base<E>& e) : Mat<eT>() { this->vec_state = 2; this->n_rows = 1; Mat<eT>::operator=(e); }
    Row& operator=(const Row& o) { Mat<eT>::operator=(static_cast<const Mat<eT>&>(o)); return *this; }
    Row& operator=(Row&& o) { Mat<eT>::operator=(static_cast<const Mat<eT>&>(o)); return *this; }
    template <class E> Row& operator=(const Base<E>& e) { Mat<eT>::operator=(e); return *this; }
};

// ------------------------------------------------------------------------------------------ Cube
template <class eT>
class Cube {
public:
    typedef eT elem_type;
    uword n_rows, n_cols, n_slices, n_elem_slice, n_elem;
    eT* mem;

private:
    std::vector<eT> own;
    bool foreign;
    mutable std::vector<std::unique_ptr<Mat<eT> > > slices;
    void make_slices() {
        slices.clear();
        for (uword s = 0; s < n_slices; ++s) slices.emplace_back(new Mat<eT>(mem + s * n_elem_slice, n_rows, n_cols, false, true));
    }
    void init_owned(uword r, uword c, uword s, eT v) {
        foreign = false;
        n_rows = r; n_cols = c; n_slices = s; n_elem_slice = r * c; n_elem = r * c * s;
        own.assign((size_t)n_elem, v);
        mem = own.data();
        make_slices();
    }

public:
    Cube() : n_rows(0), n_cols(0), n_slices(0), n_elem_slice(0), n_elem(0), mem(nullptr), foreign(false) {}
    Cube(uword r, uword c, uword s) { init_owned(r, c, s, eT(0)); }
    Cube(uword r, uword c, uword s, fill::zeros_t) { init_owned(r, c, s, eT(0)); }
    Cube(uword r, uword c, uword s, fill::ones_t) { init_owned(r, c, s, eT(1)); }
    Cube(eT* aux, uword r, uword c, uword s, bool copy_aux_mem = true, bool = false) {
        if (copy_aux_mem) { init_owned(r, c, s, eT(0)); if (n_elem) std::memcpy(mem, aux, sizeof(eT) * n_elem); }
        else { foreign = true; n_rows = r; n_cols = c; n_slices = s; n_elem_slice = r * c; n_elem = r * c * s; mem = aux; make_slices(); }
    }
    Cube(const Cube& o) { init_owned(o.n_rows, o.n_cols, o.n_slices, eT(0)); if (n_elem) std::memcpy(mem, o.mem, sizeof(eT) * n_elem); }
    Cube(const GenCube<eT>& g) { init_owned(g.r, g.c, g.s, g.v); }
    Cube& operator=(const Cube& o) {
        if (this == &o) return *this;
        if (foreign) {
            if (o.n_elem != n_elem) throw std::logic_error("refshim arma: cannot resize a cube that wraps foreign memory");
            n_rows = o.n_rows; n_cols = o.n_cols; n_slices = o.n_slices; n_elem_slice = n_rows * n_cols; make_slices();
        } else {
            init_owned(o.n_rows, o.n_cols, o.n_slices, eT(0));
        }
        if (n_elem) std::memcpy(mem, o.mem, sizeof(eT) * n_elem);
        return *this;
    }
    Cube& operator=(const GenCube<eT>& g) {
        if (foreign) {
            if (g.r * g.c * g.s != n_elem) throw std::logic_error("refshim arma: cannot resize a cube that wraps foreign memory");
            n_rows = g.r; n_cols = g.c; n_slices = g.s; n_elem_slice = n_rows * n_cols; make_slices();
            fill(g.v);
        } else {
            init_owned(g.r, g.c, g.s, g.v);
        }
        return *this;
    }
    eT& operator()(uword i, uword j, uword s) { return mem[i + j * n_rows + s * n_elem_slice]; }
    eT operator()(uword i, uword j, uword s) const { return mem[i + j * n_rows + s * n_elem_slice]; }
    eT& at(uword i, uword j, uword s) { return mem[i + j * n_rows + s * n_elem_slice]; }
    eT at(uword i, uword j, uword s) const { return mem[i + j * n_rows + s * n_elem_slice]; }
    eT& operator()(uword i) { return mem[i]; }
    eT operator()(uword i) const { return mem[i]; }
    Mat<eT>& slice(uword s) { return *slices[(size_t)s]; }
    const Mat<eT>& slice(uword s) const { return *slices[(size_t)s]; }
    Cube& fill(eT v) { for (uword i = 0; i < n_elem; ++i) mem[i] = v; return *this; }
    void zeros() { fill(eT(0)); }
    void ones() { fill(eT(1)); }
    uword size() const { return n_elem; }
    eT* memptr() { return mem; }
    const eT* memptr() const { return mem; }
};

typedef Mat<double> mat;
typedef Col<double> vec;
typedef Col<double> colvec;
typedef Row<double> rowvec;
typedef Cube<double> cube;
typedef Mat<int> imat;            // RcppArmadillo builds with 32-bit sword unless ARMA_64BIT_WORD is set
typedef Col<int> ivec;
typedef Col<int> icolvec;
typedef Row<int> irowvec;
typedef Cube<int> icube;
typedef Mat<uword> umat;
typedef Col<uword> uvec;
typedef Row<uword> urowvec;

// ---- generators
struct ZerosProxy {
    uword r, c, s; bool is_cube; double v;
    template <class eT> operator Mat<eT>() const { Mat<eT> m(r, c); m.fill((eT)v); return m; }
    template <class eT> operator Col<eT>() const { Col<eT> m(r * c); m.fill((eT)v); return m; }
    template <class eT> operator Row<eT>() const { Row<eT> m(r * c); m.fill((eT)v); return m; }
    template <class eT> operator Cube<eT>() const { Cube<eT> m(r, c, s); m.fill((eT)v); return m; }
};
inline Gen<double> zeros(uword r, uword c) { return Gen<double>(r, c, 0.0); }
inline Gen<double> ones(uword r, uword c) { return Gen<double>(r, c, 1.0); }
inline Gen<double> zeros(uword n) { return Gen<double>(n, 1, 0.0); }
inline Gen<double> ones(uword n) { return Gen<double>(n, 1, 1.0); }
inline GenCube<double> zeros(uword r, uword c, uword s) { GenCube<double> g; g.r = r; g.c = c; g.s = s; g.v = 0.0; return g; }
inline GenCube<double> ones(uword r, uword c, uword s) { GenCube<double> g; g.r = r; g.c = c; g.s = s; g.v = 1.0; return g; }
template <class T> inline T zeros(uword r, uword c) { T m(r, c); m.fill(0); return m; }
template <class T> inline T ones(uword r, uword c) { T m(r, c); m.fill(1); return m; }
template <class T> inline T zeros(uword n) { T m(n); m.fill(0); return m; }
template <class T> inline T ones(uword n) { T m(n); m.fill(1); return m; }

template <class eT> inline SizeMat size(const Mat<eT>& m) { SizeMat s; s.n_rows = m.n_rows; s.n_cols = m.n_cols; return s; }

// ---- reductions
// two interleaved accumulators, val1 + val2 at the end (arrayops::accumulate, accu_proxy_linear, op_sum proxy form)
template <class E>
inline REFSHIM_IF_EXPR(E, typename E::elem_type) accu(const E& s) {
    typedef typename E::elem_type eT;
    const uword n = s.get_n_elem();
    eT v1 = eT(0), v2 = eT(0);
    uword i, j;
    for (i = 0, j = 1; j < n; i += 2, j += 2) {
        v1 += s[i];
        v2 += s[j];
    }
    if (i < n) v1 += s[i];
    return v1 + v2;
}

// sum(): vectors -> scalar; matrices -> row vector of column sums (dim 0) or column vector of row sums (dim 1)
template <class E, int SHAPE> struct sum_impl {
    typedef typename E::elem_type result;
    static result apply(const E& s) { return accu(s); }
};
template <class E> struct sum_impl<E, 0> {
    typedef Row<typename E::elem_type> result;
    static result apply(const E& s) {
        typedef typename E::elem_type eT;
        const uword r = s.get_n_rows(), c = s.get_n_cols();
        Row<eT> out(c);
        for (uword col = 0; col < c; ++col) {
            eT v1 = eT(0), v2 = eT(0);
            uword i, j;
            for (i = 0, j = 1; j < r; i += 2, j += 2) {
                v1 += s[i + col * r];
                v2 += s[j + col * r];
            }
            if (i < r) v1 += s[i + col * r];
            out(col) = v1 + v2;
        }
        return out;
    }
};
template <class E>
inline REFSHIM_IF_EXPR(E, typename sum_impl<E, E::shape>::result) sum(const E& e) { return sum_impl<E, E::shape>::apply(e); }
template <class E>
inline REFSHIM_IF_EXPR(E, Mat<typename E::elem_type>) sum(const E& s, int dim) {
    typedef typename E::elem_type eT;
    const uword r = s.get_n_rows(), c = s.get_n_cols();
    if (dim == 0) {
        Mat<eT> out(1, c);
        for (uword col = 0; col < c; ++col) {
            eT v1 = eT(0), v2 = eT(0);
            uword i, j;
            for (i = 0, j = 1; j < r; i += 2, j += 2) { v1 += s[i + col * r]; v2 += s[j + col * r]; }
            if (i < r) v1 += s[i + col * r];
            out(0, col) = v1 + v2;
        }
        return out;
    }
    Mat<eT> out(r, 1);  // row sums: column by column, left to right
    for (uword col = 0; col < c; ++col)
        for (uword i = 0; i < r; ++i) out(i, 0) += s[i + col * r];
    return out;
}

template <class E>
inline REFSHIM_IF_EXPR(E, typename E::elem_type) max(const E& s) {
    const uword n = s.get_n_elem();
    if (!n) throw std::logic_error("max(): object has no elements");
    typename E::elem_type m = s[0];
    for (uword i = 1; i < n; ++i) { typename E::elem_type v = s[i]; if (v > m) m = v; }
    return m;
}
template <class E>
inline REFSHIM_IF_EXPR(E, typename E::elem_type) min(const E& s) {
    const uword n = s.get_n_elem();
    if (!n) throw std::logic_error("min(): object has no elements");
    typename E::elem_type m = s[0];
    for (uword i = 1; i < n; ++i) { typename E::elem_type v = s[i]; if (v < m) m = v; }
    return m;
}

inline bool is_finite(double x) { return std::isfinite(x); }
template <class E> inline REFSHIM_IF_EXPR(E, bool) is_finite(const E& s) {
    for (uword i = 0; i < s.get_n_elem(); ++i) if (!std::isfinite((double)s[i])) return false;
    return true;
}

// sort_index: std::sort over {val, index} packets, strict comparator, throws on NaN
template <class eT> struct sort_packet { eT val; uword index; };
template <class eT> struct sort_lt { bool operator()(const sort_packet<eT>& A, const sort_packet<eT>& B) const { return A.val < B.val; } };
template <class eT> struct sort_gt { bool operator()(const sort_packet<eT>& A, const sort_packet<eT>& B) const { return A.val > B.val; } };
template <class E>
inline REFSHIM_IF_EXPR(E, uvec) sort_index(const E& s, const char* dir = "ascend") {
    typedef typename E::elem_type eT;
    const uword n = s.get_n_elem();
    const char sig = dir ? dir[0] : 'a';
    if (sig != 'a' && sig != 'd') throw std::logic_error("sort_index(): unknown sort direction");
    std::vector<sort_packet<eT> > p((size_t)n);
    for (uword i = 0; i < n; ++i) {
        eT v = s[i];
        if (std::isnan((double)v)) throw std::logic_error("sort_index(): detected NaN");
        p[i].val = v; p[i].index = i;
    }
    if (sig == 'a') std::sort(p.begin(), p.end(), sort_lt<eT>());
    else std::sort(p.begin(), p.end(), sort_gt<eT>());
    uvec out(n);
    for (uword i = 0; i < n; ++i) out(i) = p[i].index;
    return out;
}

template <class E>
inline REFSHIM_IF_EXPR(E, std::ostream&) operator<<(std::ostream& os, const E& s) {
    const uword r = s.get_n_rows(), c = s.get_n_cols();
    for (uword i = 0; i < r; ++i) {
        for (uword j = 0; j < c; ++j) os << (j ? " " : "") << s[i + j * r];
        os << "\n";
    }
    return os;
}

}  // namespace arma

// =====================================================================================================
//                                               R / Rcpp
// =====================================================================================================
namespace refshim {

// The uniform stream behind Rcpp::runif / Rcpp::sample (R's unif_rand()).  Whoever drives the reference code
// installs a generator; the default throws so that an unscripted draw cannot pass unnoticed.
struct RngSource {
    virtual ~RngSource() {}
    virtual double unif_rand() = 0;                  // one uniform in (0,1)
    virtual void runif(int n, double* out) {         // Rcpp::runif(n): n uniforms, values outside (0,1) redrawn
        for (int i = 0; i < n; ++i) {
            double u;
            do { u = unif_rand(); } while (u <= 0.0 || u >= 1.0);
            out[i] = u;
        }
    }
    virtual int sample_int(int n) { return (int)(n * unif_rand() + 1); }   // Rcpp::sample(n, 1)(0), one-based
    virtual double unif_rand_for_weighted_sample() { return unif_rand(); } // the draw inside sample(x, 1, false, probs)
    virtual void set_seed(int) {}
    // position of the stream (what saving / restoring .Random.seed does in real R); generators that support it override both
    virtual long long save() { throw std::logic_error("refshim: this RngSource cannot save its state"); }
    virtual void restore(long long) { throw std::logic_error("refshim: this RngSource cannot restore its state"); }
};
inline RngSource*& rng_slot() { static thread_local RngSource* p = nullptr; return p; }
inline RngSource& rng() {
    if (!rng_slot()) throw std::logic_error("refshim: a random draw was requested but no RngSource is installed");
    return *rng_slot();
}

// R's revsort (sort a[] into descending order by heapsort, permuting ib[] alongside) — the tie order of
// Rcpp::sample's weighted branch depends on it.
inline void revsort(double* a, int* ib, int n) {
    int l, j, ir, i, ii;
    double ra;
    if (n <= 1) return;
    a--; ib--;
    l = (n >> 1) + 1;
    ir = n;
    for (;;) {
        if (l > 1) {
            l = l - 1;
            ra = a[l];
            ii = ib[l];
        } else {
            ra = a[ir];
            ii = ib[ir];
            a[ir] = a[1];
            ib[ir] = ib[1];
            if (--ir == 1) {
                a[1] = ra;
                ib[1] = ii;
                return;
            }
        }
        i = l;
        j = l << 1;
        while (j <= ir) {
            if (j < ir && a[j] > a[j + 1]) ++j;
            if (ra > a[j]) {
                a[i] = a[j];
                ib[i] = ib[j];
                j += (i = j);
            } else {
                j = ir + 1;
            }
        }
        a[i] = ra;
        ib[i] = ii;
    }
}

enum { NILSXP = 0, LGLSXP = 10, INTSXP = 13, REALSXP = 14, STRSXP = 16, VECSXP = 19, RAWSXP = 24 };

struct SexpRec;
typedef std::shared_ptr<SexpRec> SEXP;

// one R object: a typed vector with optional names / dim attributes; may wrap foreign (caller-owned) memory
struct SexpRec {
    int type;
    size_t n;
    void* data;                       // -> ints / doubles / bytes (own storage or foreign)
    std::vector<int> vi;
    std::vector<double> vd;
    std::vector<unsigned char> vr;
    std::vector<std::string> vs;
    std::vector<SEXP> vl;
    std::vector<std::string> names;
    std::vector<std::string> colnames;
    std::vector<int> dim;
    SexpRec() : type(NILSXP), n(0), data(nullptr) {}
};

inline SEXP alloc(int type, size_t n) {
    SEXP s = std::make_shared<SexpRec>();
    s->type = type; s->n = n;
    switch (type) {
        case LGLSXP: case INTSXP: s->vi.assign(n, 0); s->data = s->vi.data(); break;
        case REALSXP: s->vd.assign(n, 0.0); s->data = s->vd.data(); break;
        case RAWSXP: s->vr.assign(n, 0); s->data = s->vr.data(); break;
        case STRSXP: s->vs.assign(n, std::string()); break;
        case VECSXP: s->vl.assign(n, SEXP()); break;
        default: break;
    }
    return s;
}
inline SEXP wrap_foreign(int type, void* p, size_t n) {
    SEXP s = std::make_shared<SexpRec>();
    s->type = type; s->n = n; s->data = p;
    return s;
}
inline SEXP nil() { static SEXP s = std::make_shared<SexpRec>(); return s; }
inline bool is_nil(const SEXP& s) { return !s || s->type == NILSXP; }

inline SEXP duplicate(const SEXP& s) {
    if (is_nil(s)) return nil();
    SEXP o = alloc(s->type, s->n);
    switch (s->type) {
        case LGLSXP: case INTSXP: if (s->n) std::memcpy(o->data, s->data, s->n * sizeof(int)); break;
        case REALSXP: if (s->n) std::memcpy(o->data, s->data, s->n * sizeof(double)); break;
        case RAWSXP: if (s->n) std::memcpy(o->data, s->data, s->n); break;
        case STRSXP: o->vs = s->vs; break;
        case VECSXP: for (size_t i = 0; i < s->n; ++i) o->vl[i] = duplicate(s->vl[i]); break;
    }
    o->names = s->names; o->colnames = s->colnames; o->dim = s->dim;
    return o;
}

// coercion between INT/LGL/REAL (what R's as.integer / as.numeric would do for in-range values)
inline SEXP coerce(const SEXP& s, int type) {
    if (is_nil(s)) return alloc(type, 0);
    if (s->type == type) return s;
    if ((s->type == LGLSXP && type == INTSXP) || (s->type == INTSXP && type == LGLSXP)) {
        SEXP o = alloc(type, s->n);
        if (s->n) std::memcpy(o->data, s->data, s->n * sizeof(int));
        o->names = s->names; o->dim = s->dim; o->colnames = s->colnames;
        return o;
    }
    if ((s->type == INTSXP || s->type == LGLSXP) && type == REALSXP) {
        SEXP o = alloc(type, s->n);
        for (size_t i = 0; i < s->n; ++i) ((double*)o->data)[i] = (double)((int*)s->data)[i];
        o->names = s->names; o->dim = s->dim; o->colnames = s->colnames;
        return o;
    }
    if (s->type == REALSXP && (type == INTSXP || type == LGLSXP)) {
        SEXP o = alloc(type, s->n);
        for (size_t i = 0; i < s->n; ++i) ((int*)o->data)[i] = (int)((double*)s->data)[i];
        o->names = s->names; o->dim = s->dim; o->colnames = s->colnames;
        return o;
    }
    if (s->type == RAWSXP && type == INTSXP) {
        SEXP o = alloc(type, s->n);
        for (size_t i = 0; i < s->n; ++i) ((int*)o->data)[i] = (int)((unsigned char*)s->data)[i];
        o->dim = s->dim;
        return o;
    }
    throw std::logic_error("refshim Rcpp: unsupported coercion between R types");
}

template <int RTYPE> struct storage;
template <> struct storage<LGLSXP> { typedef int type; };
template <> struct storage<INTSXP> { typedef int type; };
template <> struct storage<REALSXP> { typedef double type; };
template <> struct storage<RAWSXP> { typedef unsigned char type; };

}  // namespace refshim

typedef refshim::SEXP SEXP;
#define R_NilValue (refshim::nil())

namespace Rcpp {

using refshim::SEXP;
using refshim::LGLSXP;
using refshim::INTSXP;
using refshim::REALSXP;
using refshim::RAWSXP;
using refshim::STRSXP;
using refshim::VECSXP;

struct exception : public std::runtime_error {
    explicit exception(const std::string& m) : std::runtime_error(m) {}
};
inline void stop(const std::string& m) { throw exception(m); }
inline void warning(const std::string& m) { std::cerr << "Warning: " << m << std::endl; }
inline void checkUserInterrupt() {}
static std::ostream& Rcout = std::cout;
static std::ostream& Rcerr = std::cerr;

class RNGScope { public: RNGScope() {} };

struct NamedPlaceHolder {};
static const NamedPlaceHolder _ = NamedPlaceHolder();

template <class T> SEXP wrap(const T& x);
template <class T> T as(const SEXP& s);

struct Named {
    std::string name;
    SEXP value;
    explicit Named(const std::string& n) : name(n) {}
    template <class T> Named(const std::string& n, const T& v) : name(n), value(wrap(v)) {}
    template <class T> Named& operator=(const T& v) { value = wrap(v); return *this; }
};

class CharacterVector;
template <int RTYPE> class Matrix;
struct ListProxy;
struct Range {
    int a, b;
    Range(int a_, int b_) : a(a_), b(b_) {}
    int size() const { return b - a + 1; }
};

// ------------------------------------------------------------------------------------------ Vector<RTYPE>
template <int RTYPE>
class Vector {
public:
    typedef typename refshim::storage<RTYPE>::type stored_type;
    typedef stored_type* iterator;
    typedef const stored_type* const_iterator;

protected:
    SEXP sx;
    stored_type* p;
    void attach(const SEXP& s) { sx = s; p = (stored_type*)sx->data; }

public:
    Vector() { attach(refshim::alloc(RTYPE, 0)); }
    Vector(const Vector& o) : sx(o.sx), p(o.p) {}   // shares, like Rcpp
    Vector(const SEXP& s) { attach(refshim::coerce(s, RTYPE)); }
    // Vector(n): n zero-initialised elements.  A negative "size" (the reference's never-evaluated default
    // arguments such as seed_vector = -1) gives an empty vector instead of an R allocation error.
    template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>
    Vector(const T& size) { attach(refshim::alloc(RTYPE, size > 0 ? (size_t)size : 0)); }
    template <class T, class U, class = typename std::enable_if<std::is_integral<T>::value && std::is_arithmetic<U>::value>::type>
    Vector(const T& size, const U& v) { attach(refshim::alloc(RTYPE, (size_t)size)); fill((stored_type)v); }
    template <class It, class = typename std::enable_if<std::is_pointer<It>::value>::type>
    Vector(It first, It last) { attach(refshim::alloc(RTYPE, (size_t)(last - first))); for (size_t i = 0; first != last; ++first, ++i) p[i] = (stored_type)*first; }
    Vector(const std::vector<int>& v) { attach(refshim::alloc(RTYPE, v.size())); for (size_t i = 0; i < v.size(); ++i) p[i] = (stored_type)v[i]; }
    Vector(std::initializer_list<stored_type> il) { attach(refshim::alloc(RTYPE, il.size())); size_t i = 0; for (auto v : il) p[i++] = v; }
    Vector& operator=(const Vector& o) { sx = o.sx; p = o.p; return *this; }
    Vector& operator=(const SEXP& s) { attach(refshim::coerce(s, RTYPE)); return *this; }
    template <class P, class = typename std::enable_if<std::is_same<P, ListProxy>::value>::type, class = void, class = void>
    Vector& operator=(const P& proxy) { return *this = proxy.get(); }
    // x = scalar: Rcpp wraps the scalar, i.e. x becomes a length-1 vector holding it
    template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>
    Vector& operator=(const T& v) { attach(refshim::alloc(RTYPE, 1)); p[0] = (stored_type)v; return *this; }
    template <class E, class = typename std::enable_if<arma::is_expr<E>::value>::type, class = void>
    Vector& operator=(const E& e) {
        attach(refshim::alloc(RTYPE, (size_t)e.get_n_elem()));
        for (size_t i = 0; i < sx->n; ++i) p[i] = (stored_type)e[i];
        return *this;
    }

    operator SEXP() const { return sx; }
    SEXP sexp() const { return sx; }
    int size() const { return (int)sx->n; }
    int length() const { return (int)sx->n; }
    stored_type& operator()(size_t i) { return p[i]; }
    const stored_type& operator()(size_t i) const { return p[i]; }
    stored_type& operator[](size_t i) { return p[i]; }
    const stored_type& operator[](size_t i) const { return p[i]; }
    stored_type& at(size_t i) { if (i >= sx->n) throw exception("index out of bounds"); return p[i]; }
    Vector operator[](const Range& r) const {   // x[Range(a, b)]: copy of the inclusive sub-range
        if (r.a < 0 || r.b >= (int)sx->n) throw exception("Range: index out of bounds");
        Vector out(r.size());
        for (int i = 0; i < r.size(); ++i) out.p[i] = p[r.a + i];
        return out;
    }
    iterator begin() { return p; }
    iterator end() { return p + sx->n; }
    const_iterator begin() const { return p; }
    const_iterator end() const { return p + sx->n; }
    Vector& fill(stored_type v) { for (size_t i = 0; i < sx->n; ++i) p[i] = v; return *this; }
    Vector& sort(bool decreasing = false) {
        if (decreasing) std::sort(p, p + sx->n, std::greater<stored_type>()); else std::sort(p, p + sx->n);
        return *this;
    }
    template <class... Args> static Vector create(Args... args) {
        const stored_type vals[] = {(stored_type)args...};
        Vector v((int)sizeof...(args));
        for (size_t i = 0; i < sizeof...(args); ++i) v.p[i] = vals[i];
        return v;
    }
    void push_back(stored_type v) {
        SEXP s = refshim::alloc(RTYPE, sx->n + 1);
        for (size_t i = 0; i < sx->n; ++i) ((stored_type*)s->data)[i] = p[i];
        ((stored_type*)s->data)[sx->n] = v;
        if (!sx->names.empty()) { s->names = sx->names; s->names.push_back(""); }
        attach(s);
    }
    void push_back(stored_type v, const std::string& name) {
        std::vector<std::string> nm = sx->names;
        nm.resize(sx->n, "");
        push_back(v);
        nm.push_back(name);
        sx->names = nm;
    }
    bool isNULL() const { return false; }
    // attr("dim") / attr("names") read access (what the Rcpp shim needs)
    std::vector<int> attr(const std::string& what) const {
        if (what == "dim") return sx->dim;
        throw exception("refshim Rcpp: attr('" + what + "') is not available");
    }
    // sugar-ish element-wise assignment from an arma vector is not needed; names:
    void names_set(const std::vector<std::string>& nm) { sx->names = nm; }
};

typedef Vector<REALSXP> NumericVector;
typedef Vector<INTSXP> IntegerVector;
typedef Vector<LGLSXP> LogicalVector;
typedef Vector<RAWSXP> RawVector;

template <int RT> inline std::ostream& operator<<(std::ostream& os, const Vector<RT>& v) {
    for (int i = 0; i < v.size(); ++i) os << (i ? " " : "") << (RT == RAWSXP ? (double)v[i] : (double)v[i]);
    return os;
}

class CharacterVector {
    SEXP sx;
public:
    CharacterVector() : sx(refshim::alloc(STRSXP, 0)) {}
    CharacterVector(const SEXP& s) : sx(refshim::is_nil(s) ? refshim::alloc(STRSXP, 0) : s) {}
    explicit CharacterVector(int n) : sx(refshim::alloc(STRSXP, (size_t)n)) {}
    template <class... Args> static CharacterVector create(Args... args) {
        const std::string vals[] = {std::string(args)...};
        CharacterVector v((int)sizeof...(args));
        for (size_t i = 0; i < sizeof...(args); ++i) v.sx->vs[i] = vals[i];
        return v;
    }
    int size() const { return (int)sx->n; }
    int length() const { return (int)sx->n; }
    std::string& operator()(size_t i) { return sx->vs[i]; }
    std::string& operator[](size_t i) { return sx->vs[i]; }
    const std::string& operator[](size_t i) const { return sx->vs[i]; }
    operator SEXP() const { return sx; }
    const std::vector<std::string>& strings() const { return sx->vs; }
};
typedef CharacterVector StringVector;

// ------------------------------------------------------------------------------------------ Matrix<RTYPE>
template <int RTYPE>
class MatrixRow {
    typedef typename refshim::storage<RTYPE>::type stored_type;
    stored_type* p0; int stride, n;
public:
    MatrixRow(stored_type* p, int stride_, int n_) : p0(p), stride(stride_), n(n_) {}
    int size() const { return n; }
    stored_type& operator[](int j) { return p0[(size_t)j * stride]; }
    stored_type& operator()(int j) { return p0[(size_t)j * stride]; }
    MatrixRow& operator=(const Vector<RTYPE>& v) {
        if (v.size() != n) throw exception("MatrixRow: size mismatch");
        for (int j = 0; j < n; ++j) p0[(size_t)j * stride] = v[j];
        return *this;
    }
    operator Vector<RTYPE>() const { Vector<RTYPE> v(n); for (int j = 0; j < n; ++j) v[j] = p0[(size_t)j * stride]; return v; }
};
template <int RTYPE>
class MatrixColumn {
    typedef typename refshim::storage<RTYPE>::type stored_type;
    stored_type* p0; int n;
public:
    MatrixColumn(stored_type* p, int n_) : p0(p), n(n_) {}
    int size() const { return n; }
    stored_type& operator[](int i) { return p0[i]; }
    stored_type& operator()(int i) { return p0[i]; }
    MatrixColumn& operator=(const Vector<RTYPE>& v) {
        if (v.size() != n) throw exception("MatrixColumn: size mismatch");
        for (int i = 0; i < n; ++i) p0[i] = v[i];
        return *this;
    }
    operator Vector<RTYPE>() const { Vector<RTYPE> v(n); for (int i = 0; i < n; ++i) v[i] = p0[i]; return v; }
};

template <int RTYPE>
class Matrix : public Vector<RTYPE> {
public:
    typedef typename Vector<RTYPE>::stored_type stored_type;
private:
    int nr, nc;
    void read_dim() {
        if (this->sx->dim.size() == 2) { nr = this->sx->dim[0]; nc = this->sx->dim[1]; }
        else { nr = (int)this->sx->n; nc = this->sx->n ? 1 : 0; }
    }
public:
    Matrix() : Vector<RTYPE>(), nr(0), nc(0) { this->sx->dim = {0, 0}; }
    Matrix(int r, int c) : Vector<RTYPE>((long)r * c), nr(r), nc(c) { this->sx->dim = {r, c}; }
    Matrix(const Matrix& o) : Vector<RTYPE>(o), nr(o.nr), nc(o.nc) {}
    Matrix(const SEXP& s) : Vector<RTYPE>(s) { read_dim(); }
    Matrix& operator=(const Matrix& o) { Vector<RTYPE>::operator=(o); nr = o.nr; nc = o.nc; return *this; }
    Matrix& operator=(const SEXP& s) { Vector<RTYPE>::operator=(s); read_dim(); return *this; }
    int nrow() const { return nr; }
    int ncol() const { return nc; }
    int rows() const { return nr; }
    int cols() const { return nc; }
    stored_type& operator()(size_t i, size_t j) { return this->p[i + j * (size_t)nr]; }
    const stored_type& operator()(size_t i, size_t j) const { return this->p[i + j * (size_t)nr]; }
    stored_type& operator()(size_t i) { return this->p[i]; }
    const stored_type& operator()(size_t i) const { return this->p[i]; }
    MatrixRow<RTYPE> row(int i) { return MatrixRow<RTYPE>(this->p + i, nr, nc); }
    MatrixColumn<RTYPE> column(int j) { return MatrixColumn<RTYPE>(this->p + (size_t)j * nr, nr); }
    MatrixRow<RTYPE> operator()(int i, NamedPlaceHolder) { return row(i); }
    MatrixColumn<RTYPE> operator()(NamedPlaceHolder, int j) { return column(j); }
};
typedef Matrix<REALSXP> NumericMatrix;
typedef Matrix<INTSXP> IntegerMatrix;
typedef Matrix<LGLSXP> LogicalMatrix;
typedef Matrix<RAWSXP> RawMatrix;

struct ColnamesProxy {
    SEXP sx;
    ColnamesProxy& operator=(const CharacterVector& cv) { sx->colnames = cv.strings(); return *this; }
};
template <int RTYPE> inline ColnamesProxy colnames(Matrix<RTYPE>& m) { ColnamesProxy p; p.sx = m.sexp(); return p; }

// ------------------------------------------------------------------------------------------ List
class List;
struct ListProxy {
    SEXP parent;
    size_t index;
    ListProxy(const SEXP& p, size_t i) : parent(p), index(i) {}
    SEXP get() const { return parent->vl[index] ? parent->vl[index] : refshim::nil(); }
    template <class T> ListProxy& operator=(const T& v) { parent->vl[index] = wrap(v); return *this; }
    ListProxy& operator=(const ListProxy& o) { parent->vl[index] = o.get(); return *this; }
    operator SEXP() const { return get(); }
    template <class T, class = typename std::enable_if<!std::is_same<T, SEXP>::value>::type>
    operator T() const { return as<T>(get()); }
};

class List {
    SEXP sx;
    void ensure_list() { if (refshim::is_nil(sx) || sx == refshim::nil()) sx = refshim::alloc(VECSXP, 0); }
public:
    List() : sx(refshim::alloc(VECSXP, 0)) {}
    explicit List(int n) : sx(refshim::alloc(VECSXP, (size_t)n)) {}
    List(const List& o) : sx(o.sx) {}
    List(const SEXP& s) : sx(s ? s : refshim::nil()) {
        if (!refshim::is_nil(sx) && sx->type != VECSXP) throw exception("refshim Rcpp: not a list");
    }
    List(const ListProxy& p) : List(p.get()) {}
    List& operator=(const List& o) { sx = o.sx; return *this; }
    List& operator=(const SEXP& s) { *this = List(s); return *this; }
    operator SEXP() const { return sx; }
    SEXP sexp() const { return sx; }
    int size() const { return refshim::is_nil(sx) ? 0 : (int)sx->n; }
    int length() const { return size(); }
    ListProxy operator[](int i) const { if (i < 0 || i >= size()) throw exception("List: index out of bounds"); return ListProxy(sx, (size_t)i); }
    ListProxy operator()(int i) const { return (*this)[i]; }
    int find(const std::string& name) const {
        if (refshim::is_nil(sx)) return -1;
        for (size_t i = 0; i < sx->names.size(); ++i) if (sx->names[i] == name) return (int)i;
        return -1;
    }
    bool containsElementNamed(const char* name) const { return find(name) >= 0; }
    ListProxy operator[](const std::string& name) const {
        int i = find(name);
        if (i < 0) throw exception("Index out of bounds: [index='" + name + "'].");
        return ListProxy(sx, (size_t)i);
    }
    ListProxy operator[](const char* name) const { return (*this)[std::string(name)]; }
    ListProxy operator()(const std::string& name) const { return (*this)[name]; }
    void push_back_sexp(const SEXP& v, const std::string* name) {
        // Rcpp's push_back builds a NEW list (other handles to the old one do not see the element)
        SEXP s = refshim::alloc(VECSXP, 0);
        if (!refshim::is_nil(sx)) { s->vl = sx->vl; s->names = sx->names; }
        bool had_names = !s->names.empty();
        s->vl.push_back(v);
        s->n = s->vl.size();
        if (name) { s->names.resize(s->n - 1, ""); s->names.push_back(*name); }
        else if (had_names) s->names.push_back("");
        sx = s;
    }
    template <class T> void push_back(const T& v) { push_back_sexp(wrap(v), nullptr); }
    template <class T> void push_back(const T& v, const std::string& name) { push_back_sexp(wrap(v), &name); }
    std::vector<std::string> names() const { return refshim::is_nil(sx) ? std::vector<std::string>() : sx->names; }

    static void add(SEXP&, size_t) {}
    template <class T, class... Rest> static void add(SEXP& s, size_t i, const T& v, const Rest&... rest) {
        put(s, i, v);
        add(s, i + 1, rest...);
    }
    template <class T> static void put(SEXP& s, size_t i, const T& v) { s->vl[i] = wrap(v); }
    static void put(SEXP& s, size_t i, const Named& v) { s->vl[i] = v.value; s->names.resize(s->n, ""); s->names[i] = v.name; }
    template <class... Args> static List create(const Args&... args) {
        List l((int)sizeof...(args));
        add(l.sx, 0, args...);
        return l;
    }
};
typedef List GenericVector;

class RObject {
    SEXP sx;
public:
    RObject() : sx(refshim::nil()) {}
    RObject(const SEXP& s) : sx(s) {}
    operator SEXP() const { return sx; }
    bool isNULL() const { return refshim::is_nil(sx); }
};

// Function / Environment: only set.seed is ever looked up (gibbs-nipt.cpp set_seed helper)
class Function {
    std::string name;
public:
    Function() {}
    explicit Function(const std::string& n) : name(n) {}
    template <class... Args> SEXP operator()(const Args&... args) const {
        if (name == "set.seed") { call_set_seed(args...); return refshim::nil(); }
        throw exception("refshim Rcpp: R function '" + name + "' is not available");
    }
private:
    template <class T, class... R> static void call_set_seed(const T& seed, const R&...) { refshim::rng().set_seed((int)seed); }
    static void call_set_seed() {}
};
class Environment {
public:
    Environment() {}
    explicit Environment(const std::string&) {}
    static Environment base_env() { return Environment(); }
    static Environment global_env() { return Environment(); }
    Function operator[](const std::string& n) const { return Function(n); }
};

// ------------------------------------------------------------------------------------------ wrap / as
namespace detail {
template <class T, class Enable = void> struct Wrap;
template <class T, class Enable = void> struct As;

template <> struct Wrap<SEXP> { static SEXP go(const SEXP& s) { return s ? s : refshim::nil(); } };
template <> struct Wrap<ListProxy> { static SEXP go(const ListProxy& p) { return p.get(); } };
template <> struct Wrap<List> { static SEXP go(const List& l) { return l.sexp(); } };
template <> struct Wrap<RObject> { static SEXP go(const RObject& l) { return (SEXP)l; } };
template <> struct Wrap<CharacterVector> { static SEXP go(const CharacterVector& l) { return (SEXP)l; } };
template <int RT> struct Wrap<Vector<RT> > { static SEXP go(const Vector<RT>& v) { return v.sexp(); } };
template <int RT> struct Wrap<Matrix<RT> > { static SEXP go(const Matrix<RT>& v) { return v.sexp(); } };
template <> struct Wrap<bool> { static SEXP go(bool b) { SEXP s = refshim::alloc(LGLSXP, 1); ((int*)s->data)[0] = b ? 1 : 0; return s; } };
template <> struct Wrap<int> { static SEXP go(int v) { SEXP s = refshim::alloc(INTSXP, 1); ((int*)s->data)[0] = v; return s; } };
template <> struct Wrap<unsigned int> { static SEXP go(unsigned int v) { SEXP s = refshim::alloc(INTSXP, 1); ((int*)s->data)[0] = (int)v; return s; } };
template <> struct Wrap<long> { static SEXP go(long v) { SEXP s = refshim::alloc(REALSXP, 1); ((double*)s->data)[0] = (double)v; return s; } };
template <> struct Wrap<unsigned long long> { static SEXP go(unsigned long long v) { SEXP s = refshim::alloc(REALSXP, 1); ((double*)s->data)[0] = (double)v; return s; } };
template <> struct Wrap<double> { static SEXP go(double v) { SEXP s = refshim::alloc(REALSXP, 1); ((double*)s->data)[0] = v; return s; } };
template <> struct Wrap<std::string> { static SEXP go(const std::string& v) { SEXP s = refshim::alloc(STRSXP, 1); s->vs[0] = v; return s; } };
template <size_t N> struct Wrap<char[N]> { static SEXP go(const char (&v)[N]) { SEXP s = refshim::alloc(STRSXP, 1); s->vs[0] = v; return s; } };
template <> struct Wrap<std::vector<int> > { static SEXP go(const std::vector<int>& v) { SEXP s = refshim::alloc(INTSXP, v.size()); if (!v.empty()) std::memcpy(s->data, v.data(), v.size() * sizeof(int)); return s; } };
template <> struct Wrap<std::vector<double> > { static SEXP go(const std::vector<double>& v) { SEXP s = refshim::alloc(REALSXP, v.size()); if (!v.empty()) std::memcpy(s->data, v.data(), v.size() * sizeof(double)); return s; } };
template <> struct Wrap<Named> { static SEXP go(const Named& n) { return n.value; } };

template <class eT> struct arma_rtype;
template <> struct arma_rtype<double> { static const int value = REALSXP; };
template <> struct arma_rtype<int> { static const int value = INTSXP; };
template <> struct arma_rtype<arma::uword> { static const int value = REALSXP; };
template <class M> inline SEXP wrap_arma_2d(const M& m) {
    typedef typename M::elem_type eT;
    SEXP s = refshim::alloc(arma_rtype<eT>::value, (size_t)m.n_elem);
    if (arma_rtype<eT>::value == REALSXP) for (arma::uword i = 0; i < m.n_elem; ++i) ((double*)s->data)[i] = (double)m[i];
    else for (arma::uword i = 0; i < m.n_elem; ++i) ((int*)s->data)[i] = (int)m[i];
    s->dim = {(int)m.n_rows, (int)m.n_cols};   // RcppArmadillo wraps Mat, Col and Row alike as an R matrix
    return s;
}
template <class eT> struct Wrap<arma::Mat<eT> > { static SEXP go(const arma::Mat<eT>& m) { return wrap_arma_2d(m); } };
template <class eT> struct Wrap<arma::Col<eT> > { static SEXP go(const arma::Col<eT>& m) { return wrap_arma_2d(m); } };
template <class eT> struct Wrap<arma::Row<eT> > { static SEXP go(const arma::Row<eT>& m) { return wrap_arma_2d(m); } };
template <class eT> struct Wrap<arma::Cube<eT> > {
    static SEXP go(const arma::Cube<eT>& m) {
        SEXP s = refshim::alloc(arma_rtype<eT>::value, (size_t)m.n_elem);
        for (arma::uword i = 0; i < m.n_elem; ++i) ((eT*)s->data)[i] = m(i);
        s->dim = {(int)m.n_rows, (int)m.n_cols, (int)m.n_slices};
        return s;
    }
};
// arma expressions are wrapped through their evaluated matrix
template <class T> struct Wrap<T, typename std::enable_if<std::is_base_of<arma::Base<T>, T>::value && !std::is_base_of<arma::Mat<typename T::elem_type>, T>::value>::type> {
    static SEXP go(const T& e) { arma::Mat<typename T::elem_type> m(e); return wrap_arma_2d(m); }
};

template <> struct As<SEXP> { static SEXP go(const SEXP& s) { return s; } };
template <> struct As<List> { static List go(const SEXP& s) { return List(s); } };
template <> struct As<CharacterVector> { static CharacterVector go(const SEXP& s) { return CharacterVector(s); } };
template <int RT> struct As<Vector<RT> > { static Vector<RT> go(const SEXP& s) { return Vector<RT>(s); } };
template <int RT> struct As<Matrix<RT> > { static Matrix<RT> go(const SEXP& s) { return Matrix<RT>(s); } };
inline double scalar_of(const SEXP& s) {
    if (refshim::is_nil(s) || s->n < 1) throw exception("Expecting a single value: [extent=0].");
    switch (s->type) {
        case LGLSXP: case INTSXP: return (double)((int*)s->data)[0];
        case REALSXP: return ((double*)s->data)[0];
        case RAWSXP: return (double)((unsigned char*)s->data)[0];
    }
    throw exception("refshim Rcpp: not a scalar");
}
template <> struct As<int> { static int go(const SEXP& s) { return (int)scalar_of(s); } };
template <> struct As<double> { static double go(const SEXP& s) { return scalar_of(s); } };
template <> struct As<bool> { static bool go(const SEXP& s) { return scalar_of(s) != 0.0; } };
template <> struct As<std::string> { static std::string go(const SEXP& s) { if (refshim::is_nil(s) || s->type != STRSXP || s->n < 1) throw exception("not a string"); return s->vs[0]; } };
template <class eT> struct As<arma::Mat<eT> > {
    static arma::Mat<eT> go(const SEXP& s0) {
        SEXP s = refshim::coerce(s0, arma_rtype<eT>::value);
        arma::uword r = s->dim.size() == 2 ? (arma::uword)s->dim[0] : (arma::uword)s->n;
        arma::uword c = s->dim.size() == 2 ? (arma::uword)s->dim[1] : 1;
        return arma::Mat<eT>((eT*)s->data, r, c, true);
    }
};
template <class eT> struct As<arma::Col<eT> > {
    static arma::Col<eT> go(const SEXP& s0) { SEXP s = refshim::coerce(s0, arma_rtype<eT>::value); return arma::Col<eT>((eT*)s->data, (arma::uword)s->n, true); }
};
template <class eT> struct As<arma::Row<eT> > {
    static arma::Row<eT> go(const SEXP& s0) { SEXP s = refshim::coerce(s0, arma_rtype<eT>::value); return arma::Row<eT>((eT*)s->data, (arma::uword)s->n, true); }
};
}  // namespace detail

template <class T> inline SEXP wrap(const T& x) { return detail::Wrap<T>::go(x); }
template <class T> inline T as(const SEXP& s) { return detail::As<T>::go(s); }
template <class T> inline T as(const ListProxy& p) { return detail::As<T>::go(p.get()); }
template <class T> inline T as(const List& l) { return detail::As<T>::go(l.sexp()); }
template <class T, int RT> inline T as(const Vector<RT>& v) { return detail::As<T>::go(v.sexp()); }
template <class T, int RT> inline T as(const Matrix<RT>& v) { return detail::As<T>::go(v.sexp()); }

template <class T> inline T clone(const T& x) { return T(refshim::duplicate((SEXP)x)); }

// ------------------------------------------------------------------------------------------ sugar
template <int RT> inline typename Vector<RT>::stored_type sum(const Vector<RT>& v) {
    typename std::conditional<RT == REALSXP, double, int>::type acc = 0;   // plain left-to-right loop
    for (int i = 0; i < v.size(); ++i) acc += v[i];
    return (typename Vector<RT>::stored_type)acc;
}
inline double max(const NumericVector& v) {
    // Rcpp sugar max: current = v[0]; returns immediately when a NaN is met
    int n = v.size();
    if (n == 0) return -std::numeric_limits<double>::infinity();
    double m = v[0];
    if (std::isnan(m)) return m;
    for (int i = 1; i < n; ++i) {
        double c = v[i];
        if (std::isnan(c)) return c;
        if (c > m) m = c;
    }
    return m;
}
inline double min(const NumericVector& v) {
    int n = v.size();
    if (n == 0) return std::numeric_limits<double>::infinity();
    double m = v[0];
    if (std::isnan(m)) return m;
    for (int i = 1; i < n; ++i) {
        double c = v[i];
        if (std::isnan(c)) return c;
        if (c < m) m = c;
    }
    return m;
}
inline int max(const IntegerVector& v) {
    int n = v.size();
    if (n == 0) return std::numeric_limits<int>::min();
    int m = v[0];
    for (int i = 1; i < n; ++i) if (v[i] > m) m = v[i];
    return m;
}
inline int min(const IntegerVector& v) {
    int n = v.size();
    if (n == 0) return std::numeric_limits<int>::max();
    int m = v[0];
    for (int i = 1; i < n; ++i) if (v[i] < m) m = v[i];
    return m;
}
// match(x, table): 1-based position of the first exact match
inline IntegerVector match(const NumericVector& x, const NumericVector& table) {
    IntegerVector out(x.size());
    for (int i = 0; i < x.size(); ++i) {
        out[i] = std::numeric_limits<int>::min();   // NA_integer_
        for (int j = 0; j < table.size(); ++j) if (table[j] == x[i]) { out[i] = j + 1; break; }
    }
    return out;
}

// ------------------------------------------------------------------------------------------ random numbers
inline NumericVector runif(int n) {
    NumericVector v(n);
    if (n > 0) refshim::rng().runif(n, v.begin());
    return v;
}
inline NumericVector runif(int n, double lo, double hi) {
    NumericVector v = runif(n);
    for (int i = 0; i < n; ++i) v[i] = lo + (hi - lo) * v[i];
    return v;
}
// sample(n, size): only size == 1 (what the path uses); one-based
inline IntegerVector sample(int n, int size, bool replace = false) {
    if (!replace && size > n) stop("Sample size must be <= n when not using replacement!");
    if (size != 1) stop("refshim Rcpp::sample(n, size): only size == 1 is implemented");
    IntegerVector out(1);
    out[0] = refshim::rng().sample_int(n);
    return out;
}
template <int RT> inline Vector<RT> sample(const Vector<RT>& x, int size, bool replace, const NumericVector& probs);
inline IntegerVector sample(int n, int size, bool replace, const NumericVector& probs, bool one_based = true) {
    IntegerVector x(n);
    for (int i = 0; i < n; ++i) x[i] = one_based ? i + 1 : i;
    return sample(x, size, replace, probs);
}
// sample(x, 1, replace, probs): Normalize, revsort, cumulative scan with one uniform
template <int RT>
inline Vector<RT> sample(const Vector<RT>& x, int size, bool replace, const NumericVector& probs) {
    const int n = x.size();
    if (probs.size() != n) stop("probs.size() != n!");
    if (size != 1) stop("refshim Rcpp::sample(x, size, replace, probs): only size == 1 is implemented");
    std::vector<double> p(probs.begin(), probs.end());
    double total = 0.0;
    int npos = 0;
    for (int i = 0; i < n; ++i) {
        if (!std::isfinite(p[i]) || p[i] < 0) stop("Probabilities must be finite and non-negative!");
        npos += (p[i] > 0.0);
        total += p[i];
    }
    if (!npos || (!replace && size > npos)) stop("Too few positive probabilities!");
    for (int i = 0; i < n; ++i) p[i] /= total;
    std::vector<int> perm((size_t)n);
    for (int i = 0; i < n; ++i) perm[i] = i + 1;
    refshim::revsort(p.data(), perm.data(), n);
    for (int i = 1; i < n; ++i) p[i] += p[i - 1];
    const double rU = refshim::rng().unif_rand_for_weighted_sample();
    int j = 0;
    for (j = 0; j < n - 1; ++j) if (rU <= p[j]) break;
    Vector<RT> out(1);
    out[0] = x[perm[j] - 1];
    return out;
}

}  // namespace Rcpp

// RcppArmadillo's container constructors from R vectors
namespace arma {
template <class eT> template <int RT> inline Mat<eT>::Mat(const Rcpp::Vector<RT>& v) : vec_state(0) {
    init_owned((uword)v.size(), 1, eT(0));
    for (uword i = 0; i < n_elem; ++i) mem[i] = (eT)v[i];
}
template <class eT> template <int RT> inline Mat<eT>::Mat(const Rcpp::Matrix<RT>& v) : vec_state(0) {
    init_owned((uword)v.nrow(), (uword)v.ncol(), eT(0));
    for (uword i = 0; i < n_elem; ++i) mem[i] = (eT)v[i];
}
template <class eT> template <int RT> inline Col<eT>::Col(const Rcpp::Vector<RT>& v) : Mat<eT>((uword)v.size(), 1) {
    this->vec_state = 1;
    for (uword i = 0; i < this->n_elem; ++i) this->mem[i] = (eT)v[i];
}
template <class eT> template <int RT> inline Row<eT>::Row(const Rcpp::Vector<RT>& v) : Mat<eT>(1, (uword)v.size()) {
    this->vec_state = 2;
    for (uword i = 0; i < this->n_elem; ++i) this->mem[i] = (eT)v[i];
}
}  // namespace arma

// R API bits the sources call directly
inline double unif_rand() { return refshim::rng().unif_rand(); }
inline int* INTEGER(const SEXP& s) { if (s->type != refshim::INTSXP && s->type != refshim::LGLSXP) throw std::logic_error("INTEGER() on a non-integer vector"); return (int*)s->data; }
inline int* LOGICAL(const SEXP& s) { return (int*)s->data; }
inline double* REAL(const SEXP& s) { if (s->type != refshim::REALSXP) throw std::logic_error("REAL() on a non-double vector"); return (double*)s->data; }
inline unsigned char* RAW(const SEXP& s) { if (s->type != refshim::RAWSXP) throw std::logic_error("RAW() on a non-raw vector"); return (unsigned char*)s->data; }
inline bool Rf_isNull(const SEXP& s) { return refshim::is_nil(s); }
inline int Rf_length(const SEXP& s) { return refshim::is_nil(s) ? 0 : (int)s->n; }
#define RcppExport extern "C"
// BEGIN_RCPP / END_RCPP: real Rcpp turns C++ exceptions into R errors; here they propagate to the test as C++ exceptions
#define BEGIN_RCPP try {
#define END_RCPP   } catch (...) { throw; }
inline void R_CheckUserInterrupt() {}
#ifndef R_NaN
#define R_NaN (std::numeric_limits<double>::quiet_NaN())
#define R_PosInf (std::numeric_limits<double>::infinity())
#define R_NegInf (-std::numeric_limits<double>::infinity())
#define NA_INTEGER (std::numeric_limits<int>::min())
#define NA_REAL (std::numeric_limits<double>::quiet_NaN())
#endif

#endif  // REFSHIM_RCPPARMADILLO_H
