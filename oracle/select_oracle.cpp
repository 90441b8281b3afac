// select_oracle.cpp — CPU ORACLE of the haplotype re-selection (test infrastructure, NOT product code).
//
// Restates, statement by statement, the in-tree R of select_new_haps_mspbwt_v3 (QUILT/R/mspbwt.R:230-474,
// heuristic_approach "A") and implements the documented contract (include/quilt_b200.h, QuiltSelectArgs) for the two
// functions of the un-vendored mspbwt 0.1.0 package it calls (map_Z_to_all_symbols, Rcpp_find_good_matches_without_a).
//
// PARITY STATUS of this file: the in-tree part follows the R source line by line (R semantics noted where they matter:
// round() is round-half-even, order() is stable, unique() keeps first occurrences); the un-vendored matching step is
// "parity unpinned" — mspbwt is not in the reference tree and cannot be installed here, so its behaviour is a contract
// re-derived from the call site (mspbwt.R:284-310), not a verified copy.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>

#include "quilt_oracle.h"

namespace {

struct Row {
    int index1, start1, end1, len1, n;  // mtm columns (mspbwt.R:310-335): 1-based haplotype, 1-based start / end in the subset, length, subset number
};

// R's round(x) for x in [0, 1] (IEC 60559 round-half-even): 0.5 -> 0
inline int r_round01(double x) { return (int)std::nearbyint(x); }

// rcpp_int_contract (STITCH): 32 SNPs per word, first SNP = least significant bit, the last word zero-padded
std::vector<uint32_t> int_contract(const double* hap_row, int stride, int nSNPs) {
    const int nG = (nSNPs + 31) / 32;
    std::vector<uint32_t> z((size_t)nG, 0u);
    for (int s = 0; s < nSNPs; s++)
        if (r_round01(hap_row[(size_t)s * stride])) z[(size_t)(s >> 5)] |= 1u << (s & 31);
    return z;
}

// map_Z_to_all_symbols (contract): the row of distinctHapsB[, g] with the same word among the rows in use, else -1
std::vector<int> map_to_symbols(const QuiltPanel* p, const std::vector<uint32_t>& z, const std::vector<int>& used) {
    std::vector<int> sym(z.size(), -1);
    for (size_t g = 0; g < z.size(); g++) {
        const int32_t* col = p->distinctHapsB + g * (size_t)p->nMaxDH;
        for (int i = 0; i < used[g]; i++)
            if ((uint32_t)col[i] == z[g]) {
                sym[g] = i + 1;
                break;
            }
    }
    return sym;
}

// Rcpp_find_good_matches_without_a (contract): rows (start0, index0, len1) of one subset
void find_good_matches(const QuiltPanel* p, const std::vector<int>& sym, const std::vector<int>& which_grids, int L, int M, std::vector<Row>& out,
                       int subset_number) {
    const int K = p->K_full, n = (int)which_grids.size();
    std::vector<int> len((size_t)K, 0);
    std::vector<int64_t> key((size_t)K);
    for (int j = 0; j < n; j++) {
        const int g = which_grids[j];
        const uint8_t* col = p->hapMatcherR + (size_t)g * K;
        const int zs = sym[g];
        for (int k = 0; k < K; k++) len[k] = (zs > 0 && col[k] == zs) ? len[k] + 1 : 0;
        // neighbours: the 2L longest current runs, ties to the lower haplotype index
        std::vector<int64_t> top;
        for (int k = 0; k < K; k++)
            if (len[k] > 0) top.push_back(((int64_t)len[k] << 32) | (uint32_t)(0x7fffffff - k));
        const size_t keep = std::min<size_t>((size_t)(2 * L), top.size());
        std::partial_sort(top.begin(), top.begin() + (long)keep, top.end(), std::greater<int64_t>());
        const bool last = (j == n - 1);
        const uint8_t* ncol = last ? nullptr : p->hapMatcherR + (size_t)which_grids[j + 1] * K;
        const int nzs = last ? -1 : sym[which_grids[j + 1]];
        for (size_t t = 0; t < keep; t++) {
            const int k = 0x7fffffff - (int)(top[t] & 0xffffffff), l = (int)(top[t] >> 32);
            const bool ends = last || !(nzs > 0 && ncol[k] == nzs);
            if (ends && l >= M) out.push_back(Row{k + 1, (j - l + 1) + 1, (j - l + 1) + l, l, subset_number});
        }
    }
}

// per haplotype of the sample: the match table `mtm` of mspbwt.R:276-350
std::vector<Row> matches_of_hap(const QuiltSelectArgs* a, int ihap, const std::vector<int>& used) {
    const QuiltPanel* p = a->panel;
    const int nGrids = p->nGrids, nIndices = a->mspbwt_nindices;
    const std::vector<uint32_t> Zs = int_contract(a->hapProbs_t + ihap, 3, p->nSNPs);
    const std::vector<int> sym = map_to_symbols(p, Zs, used);
    std::vector<Row> all;
    for (int iIndex = 1; iIndex <= nIndices; iIndex++) {
        std::vector<int> which_grids;  // seq(iIndex, nGrids, nIndices), 0-based here
        for (int g = iIndex - 1; g < nGrids; g += nIndices) which_grids.push_back(g);
        std::vector<Row> mtm;
        find_good_matches(p, sym, which_grids, a->mspbwtL, a->mspbwtM, mtm, iIndex);
        if (mtm.empty()) continue;  // return(NULL)
        if (mtm.size() > 1) {
            // mtm[order(mtm[, 1], -mtm[, "end1"], -mtm[, "start1"]), ] (stable)
            std::stable_sort(mtm.begin(), mtm.end(), [](const Row& x, const Row& y) {
                if (x.index1 != y.index1) return x.index1 < y.index1;
                if (x.end1 != y.end1) return x.end1 > y.end1;
                return x.start1 > y.start1;
            });
            // x <- c(FALSE, diff(index1) == 0 & diff(start1) == 0); mtm <- mtm[!x, ]
            std::vector<Row> kept;
            for (size_t i = 0; i < mtm.size(); i++)
                if (i == 0 || !(mtm[i].index1 == mtm[i - 1].index1 && mtm[i].start1 == mtm[i - 1].start1)) kept.push_back(mtm[i]);
            mtm.swap(kept);
        }
        all.insert(all.end(), mtm.begin(), mtm.end());  // rbind in subset order
    }
    // mtm[order(-mtm[, "len1"]), ] (stable)
    std::stable_sort(all.begin(), all.end(), [](const Row& x, const Row& y) { return x.len1 > y.len1; });
    return all;
}

std::vector<int> unique_keep_first(const std::vector<int>& v) {
    std::vector<int> out;
    std::vector<char> seen;
    for (int x : v) {
        if (x < 0) continue;  // NA
        if ((size_t)x >= seen.size()) seen.resize((size_t)x + 1, 0);
        if (!seen[(size_t)x]) {
            seen[(size_t)x] = 1;
            out.push_back(x);
        }
    }
    return out;
}

// -> the list before any padding, and the number of haplotypes found
std::vector<int> select_core(const QuiltSelectArgs* a, int* n_unique) {
    const QuiltPanel* p = a->panel;
    std::vector<int> used((size_t)p->nGrids, 0);
    for (int g = 0; g < p->nGrids; g++) {
        const uint8_t* col = p->hapMatcherR + (size_t)g * p->K_full;
        int m = 0;
        for (int k = 0; k < p->K_full; k++) m = std::max<int>(m, col[k]);
        used[(size_t)g] = std::min(m, p->nMaxDH);
    }
    std::vector<std::vector<Row>> out;
    for (int h = 0; h < a->nHap; h++) out.push_back(matches_of_hap(a, h, used));
    // unique_haps <- unique(c(out[[1]][, 1], out[[2]][, 1] (, out[[3]][, 1])))   (mspbwt.R:352-356)
    std::vector<int> cat;
    for (auto& o : out)
        for (auto& r : o) cat.push_back(r.index1);
    const std::vector<int> unique_haps = unique_keep_first(cat);
    *n_unique = (int)unique_haps.size();
    if ((int)unique_haps.size() <= a->Knew) return unique_haps;  // (:357-379; the caller pads)
    // heuristically, prioritize based on length and new-ness (:401-442)
    std::vector<std::vector<int>> results;
    size_t amax = 0;
    for (auto& mtm : out) {
        const int n = (int)mtm.size();
        std::vector<double> weight((size_t)n, 0.0);
        int m = 0;
        for (auto& r : mtm) m = std::max(m, r.end1);
        std::vector<double> cur_sum((size_t)m + 1, 1.0);
        for (int i = 0; i < n; i++) {
            const int s = mtm[i].start1, e = mtm[i].end1;
            double acc = 0;  // sum(cur_sum[s:e])
            for (int q = s; q <= e; q++) acc += cur_sum[(size_t)q];
            weight[(size_t)i] = (e - s + 1) * 1 / acc;
            for (int q = s; q <= e; q++) cur_sum[(size_t)q] += 1;
        }
        std::vector<int> o((size_t)n);
        std::iota(o.begin(), o.end(), 0);
        std::stable_sort(o.begin(), o.end(), [&](int x, int y) { return weight[(size_t)x] > weight[(size_t)y]; });  // order(-weight)
        std::vector<int> idx;
        for (int i : o) idx.push_back(mtm[(size_t)i].index1);
        amax = std::max(amax, idx.size());
        results.push_back(idx);
    }
    // pad with NA, interleave c(t(cbind(x, y (, z)))), unique, drop NA (:443-462)
    std::vector<int> inter;
    for (size_t i = 0; i < amax; i++)
        for (auto& r : results) inter.push_back(i < r.size() ? r[i] : -1);
    std::vector<int> unique_ordered_haps = unique_keep_first(inter);
    if ((int)unique_ordered_haps.size() >= a->Knew) {
        unique_ordered_haps.resize((size_t)a->Knew);
        return unique_ordered_haps;
    }
    // (:467-472) unreachable: unique_ordered_haps and unique_haps hold the same set, which is larger than Knew here
    std::vector<int> rest;
    for (int x : unique_haps)
        if (std::find(unique_ordered_haps.begin(), unique_ordered_haps.end(), x) == unique_ordered_haps.end()) rest.push_back(x);
    (void)rest;
    std::vector<int> new_haps = unique_haps;
    new_haps.resize((size_t)a->Knew);
    return new_haps;
}

bool bad(const QuiltSelectArgs* a) {
    return !a || !a->panel || !a->hapProbs_t || a->nHap < 1 || a->nHap > 3 || a->Knew < 1 || a->mspbwt_nindices < 1 || a->mspbwtL < 1 || a->mspbwtM < 1 ||
           a->panel->nGrids < a->mspbwt_nindices;
}

}  // namespace

extern "C" {

int quilt_oracle_select_haps(const QuiltSelectArgs* a, int32_t* which, int32_t* n_found, int32_t* n_unique) {
    if (bad(a) || !which || !n_found || !n_unique) return QUILT_ERR_BAD_ARG;
    int nu = 0;
    const std::vector<int> sel = select_core(a, &nu);
    *n_unique = nu;
    *n_found = (int32_t)sel.size();
    for (int i = 0; i < a->Knew; i++) which[i] = i < (int)sel.size() ? sel[(size_t)i] : 0;
    return QUILT_OK;
}

// completion of a short list as quilt_gpu_batch_chain_select does it: partial Fisher-Yates over the haplotypes not yet in
// the list, in increasing order, one uniform per draw — j = floor(n_left * u); take pool[j]; pool[j] = pool[--n_left]
int quilt_oracle_select_haps_padded(const QuiltSelectArgs* a, const double* pad_unif, int32_t* which) {
    if (bad(a) || !which || !pad_unif) return QUILT_ERR_BAD_ARG;
    int nu = 0;
    std::vector<int> sel = select_core(a, &nu);
    if ((int)sel.size() < a->Knew) {
        std::vector<char> in((size_t)a->panel->K_full + 1, 0);
        for (int x : sel) in[(size_t)x] = 1;
        std::vector<int> pool;
        for (int k = 1; k <= a->panel->K_full; k++)
            if (!in[(size_t)k]) pool.push_back(k);
        int n_left = (int)pool.size();
        int t = 0;
        while ((int)sel.size() < a->Knew && n_left > 0) {
            int j = (int)std::floor(n_left * pad_unif[t++]);
            if (j >= n_left) j = n_left - 1;
            sel.push_back(pool[(size_t)j]);
            pool[(size_t)j] = pool[(size_t)--n_left];
        }
    }
    for (int i = 0; i < a->Knew; i++) which[i] = i < (int)sel.size() ? sel[(size_t)i] : 0;
    return QUILT_OK;
}

}  // extern "C"
