// RcppEigen.h — stand-in (TEST INFRASTRUCTURE).  reference-single.cpp only uses
// Eigen::Map<Eigen::MatrixXd> as a column-major (i, j) view of R-owned memory.
#ifndef REFSHIM_RCPPEIGEN_H
#define REFSHIM_RCPPEIGEN_H
#include <cstddef>
namespace Eigen {
struct MatrixXd {};
template <class M> class Map;
template <> class Map<MatrixXd> {
    double* p;
    std::ptrdiff_t nr, nc;
public:
    Map(double* data, std::ptrdiff_t rows, std::ptrdiff_t cols) : p(data), nr(rows), nc(cols) {}
    double& operator()(std::ptrdiff_t i, std::ptrdiff_t j) { return p[i + j * nr]; }
    double operator()(std::ptrdiff_t i, std::ptrdiff_t j) const { return p[i + j * nr]; }
    std::ptrdiff_t rows() const { return nr; }
    std::ptrdiff_t cols() const { return nc; }
    double* data() { return p; }
};
}  // namespace Eigen
#endif
