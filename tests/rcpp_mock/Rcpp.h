// Minimal stand-in for <Rcpp.h> — TEST INFRASTRUCTURE ONLY.
// The build image has no R / Rcpp, so shim/quilt_gpu_shim.cpp cannot be compiled for real here.  This header declares
// just enough of the Rcpp API surface the shim uses (types, conversions, RNG helpers, the BEGIN_RCPP / END_RCPP
// brackets) for `g++ -fsyntax-only` to type-check the translation unit (tests/test_cabi_host.py).  Nothing here is
// linked or executed; semantics are the real Rcpp's.
#pragma once
#include <cstddef>
#include <stdexcept>
#include <string>
#include <vector>

typedef struct SEXPREC* SEXP;
int* INTEGER(SEXP);
double* REAL(SEXP);
unsigned char* RAW(SEXP);
#define RcppExport extern "C"
#define BEGIN_RCPP try {
#define END_RCPP                  \
    }                             \
    catch (std::exception&) {     \
        return (SEXP)0;           \
    }                             \
    return (SEXP)0;

namespace Rcpp {

class RObject {
   public:
    RObject() {}
    RObject(SEXP) {}
    operator SEXP() const { return (SEXP)0; }
    struct AttrProxy {
        template <class T>
        operator T() const { return T(); }
    };
    AttrProxy attr(const char*) const { return AttrProxy(); }
};

template <class T>
class Vector : public RObject {
   public:
    Vector() {}
    Vector(SEXP) {}
    explicit Vector(int) {}
    int size() const { return 0; }
    T& operator[](int) { return v_; }
    const T& operator[](int) const { return v_; }
    T& operator()(int) { return v_; }
    const T& operator()(int) const { return v_; }
    T* begin() { return &v_; }
    T* end() { return &v_; }
    const T* begin() const { return &v_; }
    const T* end() const { return &v_; }

   private:
    T v_ = T();
};
typedef Vector<int> IntegerVector;
typedef Vector<double> NumericVector;
typedef Vector<int> LogicalVector;

class CharacterVector : public RObject {
   public:
    template <class... A>
    static CharacterVector create(A...) { return CharacterVector(); }
};

template <class T>
class Matrix : public RObject {
   public:
    Matrix() {}
    Matrix(SEXP) {}
    Matrix(int, int) {}
    int nrow() const { return 0; }
    int ncol() const { return 0; }
};
typedef Matrix<int> IntegerMatrix;
typedef Matrix<double> NumericMatrix;
typedef Matrix<unsigned char> RawMatrix;

class List : public RObject {
   public:
    List() {}
    List(SEXP) {}
    explicit List(int) {}
    int size() const { return 0; }
    struct Proxy {
        template <class T>
        operator T() const { return T(); }
        template <class T>
        Proxy& operator=(const T&) { return *this; }
    };
    Proxy operator[](int) const { return Proxy(); }
    Proxy operator[](const char*) const { return Proxy(); }
    template <class T>
    void push_back(const T&, const char*) {}
};

template <class T>
T as(SEXP) { return T(); }
template <class T>
T as(const List::Proxy&) { return T(); }
template <class T>
T clone(const T& x) { return x; }

NumericVector runif(int n);
IntegerVector sample(int n, int size);
[[noreturn]] inline void stop(const std::string& msg) { throw std::runtime_error(msg); }
class RNGScope {};

struct ColnamesProxy {
    ColnamesProxy& operator=(const CharacterVector&) { return *this; }
};
template <class T>
ColnamesProxy colnames(Matrix<T>&) { return ColnamesProxy(); }

}  // namespace Rcpp

// REAL / INTEGER on Rcpp containers (implicit SEXP conversion in real Rcpp)
