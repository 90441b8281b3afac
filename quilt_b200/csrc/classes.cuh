// classes.cuh — haplotype classes of a grid: the K selected haplotypes grouped by the allele bits the grid's reads can see.
//
// Every read of grid g takes its emission value E[k, r] from the allele pattern of haplotype k at the read's SNPs
// (reference: eMatRead_t, gibbs-small.cpp:116-265).  Those SNPs lie in the grid's own 32-SNP word and — for reads that start
// before / end after it — in a few bits of the neighbouring words.  Two haplotypes with the same own word and the same
// neighbour bits are indistinguishable to every read of the grid: inside sample_reads_in_grid (gibbs-nipt.cpp:733-1295) they
// always receive the same emission factor.  With C such classes (a few hundred at most on a compressed panel: the
// distinct-haplotype idea of QUILT2's own hapMatcherR, here on the selected subset and including the read-visible neighbour
// bits) the K-long sums of the read resampler collapse to C-long sums over class totals of alphaHat_m * betaHat_m, kept one
// class per thread by k_sweep.  This kernel builds, once per call and grid:
//   class ids (rank of the class key), the haplotypes sorted by (class, k) for the segmented class sums, the
//   thread-major class ids for the lazy column update, the class records {words, offset | count} and cinfo[g] = C (or 0).
// Everything is a deterministic function of (W, read descriptors): ids never depend on thread timing.
#pragma once

#include "device_common.cuh"
#include "types.h"

namespace qb {

constexpr int CLS_HS = 1024;  // hash slots (load <= 25 %)
constexpr int CLS_NT = 512;   // threads of k_build_classes
constexpr int CLS_NW = CLS_NT / 32;

struct ClassSmem {
    int owner[CLS_HS];      // 0 = empty, else k + 1 of the haplotype that claimed the slot (its key is the slot's key)
    uint16_t scls[CLS_HS];  // slot -> class id
    uint16_t list[256];     // occupied slots
    int ccount[256];        // members per class
    int coffs[257];         // first sorted position per class
    int crep[256];          // representative haplotype per class
    uint16_t cntW[CLS_NW][256];  // per-warp members per class (counting sort)
    uint4 crk[256];         // occupied slots, compact: {members, key words} for the ranking loop
    int wtot[8];            // class-size totals of the eight 32-class groups
    uint32_t mp, mn;        // neighbour bits the grid's reads can see (previous / next word)
    int impure, overflow, ndistinct, nlist;
};

// dynamic shared memory: key[3][Kp] u32, slot_of / local position [Kp] u16, class [Kp] u8, sorted [KA] u16
__host__ __device__ inline size_t class_dyn_smem(int Kp, int KA) { return (size_t)Kp * 12 + (size_t)Kp * 2 + (size_t)((Kp + 15) & ~15) + (size_t)KA * 2 + 16; }

// grid = (T, jobs), CLS_NT threads.  NT / EPT: geometry of the sweep kernel that will consume the layouts (KA = NT * EPT >= K).
__global__ void __launch_bounds__(CLS_NT) k_build_classes(BatchParams P, const JobDev* __restrict__ jobs, int NT, int EPT) {
    extern __shared__ __align__(16) unsigned char cls_dyn[];
    __shared__ ClassSmem S;
    const JobDev& J = jobs[blockIdx.y];
    const int g = blockIdx.x, K = P.K, Kp = P.Kp, T = P.T, KA = NT * EPT;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t* key0 = reinterpret_cast<uint32_t*>(cls_dyn);
    uint32_t* key1 = key0 + Kp;
    uint32_t* key2 = key1 + Kp;
    uint16_t* slot_of = reinterpret_cast<uint16_t*>(key2 + Kp);
    uint8_t* clsk = reinterpret_cast<uint8_t*>(slot_of + Kp);
    uint16_t* sorted = reinterpret_cast<uint16_t*>(clsk + ((Kp + 15) & ~15));
    const int r0 = J.rs[g], r1 = J.rs[g + 1];
    const int n_g = r1 - r0;
    for (int i = tid; i < CLS_HS; i += CLS_NT) S.owner[i] = 0;
    if (tid < 256) S.ccount[tid] = 0;
    for (int i = tid; i < CLS_NW * 256; i += CLS_NT) (&S.cntW[0][0])[i] = 0;
    if (tid == 0) {
        S.mp = 0;
        S.mn = 0;
        S.overflow = 0;
        S.ndistinct = 0;
        S.nlist = 0;
        // only grids whose reads arrive in ONE staging chunk (all but over-full grids) take the class path
        S.impure = (n_g <= 0 || J.ginfo[4 * g + 2] != n_g) ? 1 : 0;
    }
    __syncthreads();
    // ---- which neighbour bits can the grid's reads see?  (table-mode reads on consecutive SNPs only)
    for (int r = r0 + tid; r < r1; r += CLS_NT) {
        const ReadDesc d = J.desc[r];
        if (d.mode != MODE_RUN) {
            S.impure = 1;
            continue;
        }
        const uint64_t run = (((uint64_t)1 << d.nb) - 1) << d.b0;  // bits of the run relative to its first word
        const uint32_t lo = (uint32_t)run, hi = (uint32_t)(run >> 32);
        if (d.g0rel < 0) {
            atomicOr(&S.mp, lo);  // (hi lands in the grid's own word)
        } else if (d.g0rel == 0) {
            if (hi) atomicOr(&S.mn, hi);
        } else {
            atomicOr(&S.mn, lo);  // (the host keeps reads reaching beyond word g + 1 as dense columns)
        }
    }
    __syncthreads();
    const bool pure = !S.impure;
    if (pure) {
        const uint32_t mp = (g > 0) ? S.mp : 0u, mn = (g + 1 < T) ? S.mn : 0u;
        // ---- keys + hash insertion (the slot's key is the key of the haplotype that claimed it)
        for (int k = tid; k < K; k += CLS_NT) {
            key0[k] = J.W[(size_t)g * Kp + k];
            key1[k] = mp ? (J.W[(size_t)(g - 1) * Kp + k] & mp) : 0u;
            key2[k] = mn ? (J.W[(size_t)(g + 1) * Kp + k] & mn) : 0u;
        }
        __syncthreads();
        const bool nb_keys = (mp | mn) != 0;  // (most grids: no read sees a neighbour bit, the own word is the whole key)
        for (int k = tid; k < K; k += CLS_NT) {
            const uint32_t a = key0[k], b = key1[k], c = key2[k];
            uint32_t h = (a * 0x9E3779B1u) ^ (b * 0x85EBCA6Bu) ^ (c * 0xC2B2AE35u);
            h ^= h >> 15;
            int s = (int)(h & (CLS_HS - 1));
            bool placed = false;
            while (!*(volatile int*)&S.overflow) {
                // (nearly every haplotype finds its class already in the table: look before trying to claim)
                int o = *(volatile int*)&S.owner[s];
                if (o == 0) {
                    o = atomicCAS(&S.owner[s], 0, k + 1);
                    if (o == 0) {
                        if (atomicAdd(&S.ndistinct, 1) >= CLS_MAX) S.overflow = 1;
                        placed = true;
                        break;
                    }
                }
                // (the owner's key words were written before the barrier above: plain reads)
                if (key0[o - 1] == a && (!nb_keys || (key1[o - 1] == b && key2[o - 1] == c))) {
                    placed = true;
                    break;
                }
                s = (s + 1) & (CLS_HS - 1);
            }
            if (placed) slot_of[k] = (uint16_t)s;
        }
    }
    __syncthreads();
    const bool ok = pure && !S.overflow;
    if (!ok) {
        if (tid == 0) J.cinfo[g] = 0;
        return;
    }
    // ---- occupied slots -> classes ranked by (members descending, key ascending)
    for (int s = tid; s < CLS_HS; s += CLS_NT)
        if (S.owner[s]) S.list[atomicAdd(&S.nlist, 1)] = (uint16_t)s;
    __syncthreads();
    const int nC = S.nlist;
    if (tid < nC) {
        const int s = S.list[tid], ko = S.owner[s] - 1;
        S.crk[tid] = make_uint4(0u, key0[ko], key1[ko], key2[ko]);
    }
    __syncthreads();
    {
        // class id = rank of the key (own word, previous-word bits, next-word bits): CLS_NT / 256 threads per class share the
        // comparisons (integer counts: the order of the atomic adds does not matter)
        constexpr int PER = CLS_NT / 256;
        const int i = tid & 255, part = tid >> 8;
        if (i < nC) {
            const uint4 me = S.crk[i];
            const int q0 = (nC * part) / PER, q1 = (nC * (part + 1)) / PER;
            int rank = 0;
            for (int q = q0; q < q1; q++) {
                const uint4 o = S.crk[q];
                rank += ((o.y < me.y) || (o.y == me.y && (o.z < me.z || (o.z == me.z && o.w < me.w)))) ? 1 : 0;
            }
            atomicAdd(&S.ccount[i], rank);  // (ccount is the scratch of the ranking here; it takes the class sizes below)
        }
    }
    __syncthreads();
    if (tid < nC) {
        const int s = S.list[tid], rank = S.ccount[tid];
        S.scls[s] = (uint16_t)rank;
        // representative: any member gives the same key; take the slot's owner (which haplotype claimed the slot depends on
        // timing, its key does not)
        S.crep[rank] = S.owner[s] - 1;
    }
    __syncthreads();
    for (int k = tid; k < K; k += CLS_NT) clsk[k] = (uint8_t)S.scls[slot_of[k]];
    __syncthreads();
    // ---- stable counting sort by class: warp w owns the haplotypes [w * Kw, w * Kw + Kw)
    const int Kw = (((K + CLS_NW - 1) / CLS_NW) + 31) & ~31;
    for (int st = 0; st < Kw; st += 32) {
        const int k = warp * Kw + st + lane;
        const bool in = k < K;
        const int c = in ? (int)clsk[k] : -1;
        const unsigned peers = __match_any_sync(0xffffffffu, c);
        const int before = __popc(peers & ((1u << lane) - 1u));
        int base = 0;
        if (in) base = S.cntW[warp][c];
        __syncwarp();
        if (in) {
            slot_of[k] = (uint16_t)(base + before);                                   // position among the warp's members of the class
            if (before == __popc(peers) - 1) S.cntW[warp][c] = (uint16_t)(base + before + 1);  // (the last peer holds the new count)
        }
        __syncwarp();
    }
    __syncthreads();
    // class sizes = the warps' counts; first sorted position per class = exclusive prefix over the class ids
    if (tid < nC) {
        int n = 0;
        for (int w = 0; w < CLS_NW; w++) n += S.cntW[w][tid];
        S.ccount[tid] = n;
    }
    __syncthreads();
    {
        // exclusive prefix over the class ids: warp scans of 32 + the warps' totals
        int v = (tid < nC) ? S.ccount[tid] : 0, incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += y;
        }
        if (tid < 256 && lane == 31) S.wtot[warp] = incl;
        __syncthreads();
        if (tid <= nC && tid < 257) {
            int o = 0;
            for (int w = 0; w < (tid >> 5); w++) o += S.wtot[w];
            S.coffs[tid] = (tid < 256) ? o + incl - v : o;
        }
    }
    __syncthreads();
    if (tid < nC) {
        int o = S.coffs[tid];
        for (int w = 0; w < CLS_NW; w++) {
            const int n = S.cntW[w][tid];
            S.cntW[w][tid] = (uint16_t)o;
            o += n;
        }
    }
    __syncthreads();
    for (int st = 0; st < Kw; st += 32) {
        const int k = warp * Kw + st + lane;
        if (k < K) {
            const int c = clsk[k];
            const int pos = (int)S.cntW[warp][c] + (int)slot_of[k];
            sorted[pos] = (uint16_t)(k | ((pos == S.coffs[c]) ? 0x8000 : 0));
        }
    }
    for (int j = K + tid; j < KA; j += CLS_NT) sorted[j] = (uint16_t)(j | ((j == K) ? 0x8000 : 0));  // padding elements: one class of zeros
    __syncthreads();
    // ---- outputs
    uint16_t* operm = J.cperm + (size_t)g * KA;
    for (int j = tid; j < KA; j += CLS_NT) operm[j] = sorted[j];
    uint8_t* ocls = J.ccls + (size_t)g * KA;
    for (int q = tid; q < KA; q += CLS_NT) {
        const int t = q / EPT, i = q - t * EPT;
        const int k = t + i * NT;
        ocls[q] = (k < K) ? clsk[k] : (uint8_t)nC;
    }
    uint8_t* oent = J.cent + (size_t)g * NT;
    for (int t = tid; t < NT; t += CLS_NT) {
        int e = 0;
        if (t > 0) {
            const int p = EPT * t - 1;
            e = 1 + ((p < K) ? (int)clsk[sorted[p] & 0x7fff] : nC);
        }
        oent[t] = (uint8_t)e;
    }
    uint4* orec = J.crec + (size_t)g * (CLS_LANES * 32);
    for (int c = tid; c < CLS_LANES * 32; c += CLS_NT) {
        uint4 rec = make_uint4(0u, 0u, 0u, 0u);
        if (c < nC) {
            const int kr = S.crep[c];
            rec = make_uint4(key1[kr], key0[kr], key2[kr], (uint32_t)S.coffs[c] | ((uint32_t)S.ccount[c] << 16));
        }
        orec[c] = rec;
    }
    if (tid == 0) J.cinfo[g] = nC;
}

}  // namespace qb
