/* quilt_oracle.h — entry points of the CPU oracle (oracle/libquiltoracle.so).  TEST INFRASTRUCTURE ONLY: a statement-order
 * restatement of the reference C++, same flat structs as the product ABI (include/quilt_b200.h).  Nothing under
 * quilt_b200/ may include this file. */
#ifndef QUILT_ORACLE_H
#define QUILT_ORACLE_H
#include "../include/quilt_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
int quilt_oracle_gibbs(const QuiltGibbsArgs* args, QuiltGibbsOut* out);
int quilt_oracle_make_eMatRead_t(const QuiltGibbsArgs* args, double* eMatRead_t, int32_t* read_category);
int quilt_oracle_unpack_panel(const QuiltPanel* panel, int32_t K, const int32_t* which_haps_to_use, int32_t all_snps, uint32_t* words);
int quilt_oracle_forward_backward(int32_t K, int32_t nGrids, const double* eMatGrid_t, const double* transMatRate_tc_H,
                                  double* alphaHat_t, double* betaHat_t, double* c);
/* select_new_haps_mspbwt_v3 (QUILT/R/mspbwt.R:230-474), contract in include/quilt_b200.h */
int quilt_oracle_select_haps(const QuiltSelectArgs* args, int32_t* which_haps_to_use, int32_t* n_found, int32_t* n_unique);
/* the same with the short-list completion of quilt_gpu_batch_chain_select (pad_unif [Knew], prev_unused: ignored) */
int quilt_oracle_select_haps_padded(const QuiltSelectArgs* args, const double* pad_unif, int32_t* which_haps_to_use);
#ifdef __cplusplus
}
#endif
#endif
