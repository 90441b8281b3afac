// device_common.cuh — shared device helpers for the sm_100a kernels.
//
// Numerical contract (DESIGN.md §"arithmetic"): every element-wise operation that lands in
// stored state (alpha, beta, eMatGrid, c, emission tables) is the same IEEE fp64 operation, in
// the same association, as the reference's Armadillo expression; the translation unit is built
// with -fmad=false so nvcc never contracts a*b+c into an FMA (the reference is baseline x86-64,
// no FMA).  Only the order of K-long sums differs (tree instead of the CPU's two accumulators).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#define QB_WARP 32

namespace qb {

// ---------------------------------------------------------------- block-wide sum of V doubles
// Every thread returns the same, bitwise-identical totals: xor-butterfly inside the warp (all lanes
// end with the same bits), one shared-memory slot per warp, then every thread adds the warp
// partials in warp order.  Two alternating scratch buffers make one __syncthreads per call enough.
template <int NT, int V>
struct BlockSum {
    static constexpr int NW = NT / QB_WARP;
    double* scratch;  // [2][V][NW]
    int phase;
    __device__ __forceinline__ BlockSum(double* s) : scratch(s), phase(0) {}
    static constexpr int scratch_doubles() { return 2 * V * NW; }

    __device__ __forceinline__ void run(double (&v)[V]) {
#pragma unroll
        for (int i = 0; i < V; i++) {
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], d);
        }
        if (NW == 1) return;
        double* buf = scratch + phase * (V * NW);
        phase ^= 1;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < V; i++) buf[i * NW + warp] = v[i];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < V; i++) {
            double s = buf[i * NW];
#pragma unroll
            for (int w = 1; w < NW; w++) s += buf[i * NW + w];
            v[i] = s;
        }
    }
};

// max over the block (used by the emission build; order-free, so bit-exact vs the CPU)
template <int NT>
__device__ __forceinline__ double block_max(double v, double* scratch /*[NT/32]*/) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, d));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double m = scratch[0];
#pragma unroll
    for (int w = 1; w < NT / 32; w++) m = fmax(m, scratch[w]);
    return m;
}

// ---------------------------------------------------------------- 1-D bulk async copy (TMA engine) + mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// spin with a watchdog: a lost arrival traps (-> CUDA error on the host) instead of hanging the device
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();
    }
}
// global -> shared::cta bulk copy, completion signalled on the mbarrier (bytes multiple of 16, 16-B aligned)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------- thread-block cluster (K > 4096: two CTAs per job)
// shared -> global bulk copy (TMA engine, bulk async-group completion): the issuing thread commits a group and later waits either
// for the source to have been read (the shared-memory buffer may be overwritten) or for the whole group (the writes are visible)
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
// L2 prefetch of a whole range by the TMA engine: one instruction instead of one prefetch.global.L2 per 128-byte line (which
// fills the LSU queue of the warps that issue them).  Address 16-byte aligned, size a positive multiple of 16.
__device__ __forceinline__ void bulk_prefetch_l2(const void* gptr, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}
// the same for an arbitrary byte range inside an allocation: start rounded down, end rounded down to 16 bytes
__device__ __forceinline__ void bulk_prefetch_l2_range(const void* p, size_t bytes) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p), lo = a & ~(uintptr_t)15, hi = (a + bytes) & ~(uintptr_t)15;
    if (hi > lo) bulk_prefetch_l2(reinterpret_cast<const void*>(lo), (uint32_t)(hi - lo));
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// store a double into the same shared-memory location of CTA `rank` of the cluster (distributed shared memory)
__device__ __forceinline__ void st_dsmem(double* local_ptr, uint32_t rank, double v) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local_ptr)), "r"(rank));
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(ra), "d"(v) : "memory");
}

// streaming global access (state columns are touched once per pass: keep them out of L1)
__device__ __forceinline__ double ld_stream(const double* p) {
    double v;
    asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void st_stream(double* p, double v) { asm volatile("st.global.L1::no_allocate.f64 [%0], %1;" ::"l"(p), "d"(v)); }

}  // namespace qb

namespace qb {

// ---------------------------------------------------------------- small irregular staging: cp.async (LDGSTS)
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// generic-proxy writes (global or shared) -> later async-proxy (bulk copy) reads
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ double ld_cg(const double* p) {
    double v;
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

// per-thread element mapping of a K-long column: k = tid + i * NT (coalesced), masked by k < K
template <int NT, int EPT>
struct Col {
    __device__ static __forceinline__ void load(double (&v)[EPT], const double* __restrict__ p, int K, double fill) {
#pragma unroll
        for (int i = 0; i < EPT; i++) {
            const int k = threadIdx.x + i * NT;
            v[i] = (k < K) ? ld_stream(p + k) : fill;
        }
    }
    __device__ static __forceinline__ void load_smem(double (&v)[EPT], const double* p, int K, double fill) {
#pragma unroll
        for (int i = 0; i < EPT; i++) {
            const int k = threadIdx.x + i * NT;
            v[i] = (k < K) ? p[k] : fill;
        }
    }
    __device__ static __forceinline__ void store(const double (&v)[EPT], double* __restrict__ p, int K) {
#pragma unroll
        for (int i = 0; i < EPT; i++) {
            const int k = threadIdx.x + i * NT;
            if (k < K) st_stream(p + k, v[i]);
        }
    }
    // this thread's share of a K-long sum as a pairwise tree: log2(EPT) dependent additions instead of EPT - 1 (the block-wide sums
    // are trees over lanes and warps anyway, so no summation order of the reference is given up here)
    __device__ static __forceinline__ double sum(const double (&v)[EPT]) {
        double t[EPT];
#pragma unroll
        for (int i = 0; i < EPT; i++) t[i] = v[i];
#pragma unroll
        for (int st = 1; st < EPT; st <<= 1) {
#pragma unroll
            for (int i = 0; i + st < EPT; i += 2 * st) t[i] += t[i + st];
        }
        return t[0];
    }
};

}  // namespace qb
