// quilt_gpu.cu — host side of libquiltgpu.so: the C ABI of include/quilt_b200.h over the sm_100a kernels.
//
// Work unit ("job") = one rcpp_forwardBackwardGibbsNIPT call (QUILT/src/gibbs-nipt.cpp:2395-3307).  Jobs of a
// batch are grouped into buckets of identical shape/flags; each bucket runs in waves of at most
// (SM count x resident CTAs per SM) jobs, one CTA per job in the sweep kernel.  The K x T state (alpha, beta,
// eMatGrid, allele words, emission tables) belongs to a wave slot and is reused by the next wave; only the
// small per-job inputs and outputs stay resident for the whole batch.
//
// There is no CPU fallback anywhere in this file: without a CUDA device every entry point returns
// QUILT_ERR_NO_DEVICE / QUILT_ERR_CUDA.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#include "../../include/quilt_b200.h"
#include "block_nipt.cuh"
#include "passes.cuh"
#include "select.cuh"
#include "haploid.cuh"
#include "prep.cuh"
#include "classes.cuh"
#include "io_rows.cuh"
#include "sweep.cuh"
#include "types.h"

using namespace qb;

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};
std::mutex g_mu;
cudaStream_t g_stream = nullptr;
cudaStream_t g_copy_in = nullptr, g_copy_out = nullptr;  // staging streams of the pipelined host-buffer path
int g_device = -1;
int g_sms = 0;

int set_err(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess)                                                                           \
            return set_err(QUILT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__) + " @" + std::to_string(__LINE__)); \
    } while (0)
#define LAUNCHED() (g_launches.fetch_add(1, std::memory_order_relaxed))

// Per-section timing — the reference's print_extra_timing_information / suppressOutput = 0 switch (copied-from-stitch.cpp:31-45
// prints the wall time between code sections): when enabled, every tagged launch site records a CUDA event behind its kernel.  The
// library stream is in-order, so the time between two consecutive events is the duration of the kernel launched in between (plus
// its launch gap; copies queued between two waves are charged to the first kernel after them).  Off by default: no events, no cost.
struct SectionTimer {
    bool on = false;
    cudaEvent_t begin = nullptr;
    std::vector<std::pair<const char*, cudaEvent_t>> marks;
    std::mutex mu;
    void reset() {
        for (auto& m : marks) cudaEventDestroy(m.second);
        marks.clear();
        if (begin) cudaEventDestroy(begin);
        begin = nullptr;
    }
};
SectionTimer g_sect;
// QUILT_B200_TRACE=1: host wall-clock stamps of the pipelined paths on stderr (where the end-to-end time outside the kernels goes)
inline bool trace_on() {
    static const bool on = std::getenv("QUILT_B200_TRACE") != nullptr;
    return on;
}
inline double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
inline void launched_k(const char* name) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (!g_sect.on) return;
    std::lock_guard<std::mutex> lk(g_sect.mu);
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, g_stream);
    g_sect.marks.emplace_back(name, e);
}
#define LAUNCHED_K(name) launched_k(name)

int ensure_device() {
    if (g_stream) return QUILT_OK;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return set_err(QUILT_ERR_NO_DEVICE, "no CUDA device: libquiltgpu has no CPU fallback");
    if (g_device < 0) g_device = 0;
    CK(cudaSetDevice(g_device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, g_device));
    if (prop.major < 10) return set_err(QUILT_ERR_NO_DEVICE, "libquiltgpu is built for sm_100a only");
    g_sms = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&g_copy_in, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&g_copy_out, cudaStreamNonBlocking));
    return QUILT_OK;
}

inline size_t al(size_t x) { return (x + 255) & ~(size_t)255; }

struct DBuf {
    void* p = nullptr;
    size_t bytes = 0;
    DBuf() {}
    DBuf(const DBuf&) = delete;
    DBuf& operator=(const DBuf&) = delete;
    ~DBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    cudaError_t alloc(size_t n) {
        release();
        if (n == 0) n = 256;
        cudaError_t e = cudaMalloc(&p, n);
        if (e == cudaSuccess) bytes = n;
        return e;
    }
};
struct HBuf {  // pinned host
    void* p = nullptr;
    size_t bytes = 0;
    HBuf() {}
    HBuf(const HBuf&) = delete;
    HBuf& operator=(const HBuf&) = delete;
    ~HBuf() { release(); }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        bytes = 0;
    }
    cudaError_t alloc(size_t n) {
        release();
        if (n == 0) n = 256;
        cudaError_t e = cudaMallocHost(&p, n);
        if (e == cudaSuccess) bytes = n;
        return e;
    }
};

// Staging buffers are expensive to create (cudaMallocHost / cudaMalloc of several GB): batches hand them back to a
// small cache on quilt_gpu_batch_free and the next batch reuses them.
template <class B>
struct BufCache {
    std::vector<std::unique_ptr<B>> free_;
    std::unique_ptr<B> acquire(size_t n, cudaError_t* err) {
        *err = cudaSuccess;
        int best = -1;
        for (int i = 0; i < (int)free_.size(); i++)
            if (free_[i]->bytes >= n && (best < 0 || free_[i]->bytes < free_[best]->bytes)) best = i;
        if (best >= 0) {
            std::unique_ptr<B> b = std::move(free_[best]);
            free_.erase(free_.begin() + best);
            return b;
        }
        // nothing fits: grow.  Several batches are alive at once in a call chain, so the other cached buffers stay (they fit the
        // other stages); only when the cache is full does the smallest one go.
        if (free_.size() >= CAP) {
            int small = 0;
            for (int i = 1; i < (int)free_.size(); i++)
                if (free_[i]->bytes < free_[small]->bytes) small = i;
            free_.erase(free_.begin() + small);
        }
        std::unique_ptr<B> b(new B());
        *err = b->alloc(n);
        if (*err != cudaSuccess) {
            free_.clear();  // out of memory: drop everything cached and try once more
            *err = b->alloc(n);
        }
        return b;
    }
    static constexpr size_t CAP = 8;
    void release(std::unique_ptr<B> b) {
        if (b && b->p) free_.push_back(std::move(b));
        while (free_.size() > CAP) free_.erase(free_.begin());
    }
    void clear() { free_.clear(); }
};
BufCache<DBuf> g_dcache_in, g_dcache_out;
// small per-batch buffers (device-only result zone, JobDev records): cudaMalloc / cudaFree / cudaMallocHost per batch cost
// milliseconds each and cudaFree synchronises the device, so they are recycled too
BufCache<DBuf> g_dcache_outB, g_dcache_jobs, g_dcache_hap;
template <class B>
cudaError_t cached_alloc(B& b, BufCache<B>& cache, size_t n) {
    cudaError_t e;
    std::unique_ptr<B> t = cache.acquire(n == 0 ? 256 : n, &e);
    if (e != cudaSuccess) return e;
    b.release();
    std::swap(b.p, t->p);
    std::swap(b.bytes, t->bytes);
    return cudaSuccess;
}
template <class B>
void cached_release(B& b, BufCache<B>& cache) {
    if (!b.p) return;
    std::unique_ptr<B> t(new B());
    std::swap(b.p, t->p);
    std::swap(b.bytes, t->bytes);
    cache.release(std::move(t));
}
// The wave state (alpha / beta / eMatGrid / allele words / tables of the jobs in flight) is scratch that only lives while
// a batch runs.  Batches run one after the other on the library stream, so ALL staged batches share one arena (grown to
// the largest request): several batches can be staged at once — the stages of a device-resident call chain — without
// multiplying tens of GB of state.  JobDev records are rebuilt at every run (upload_jobdevs), so the arena may move.
DBuf g_slots;
BufCache<HBuf> g_hcache_in, g_hcache_out, g_hcache_jobs;

int host_threads() {
    static int n = 0;
    if (n == 0) {
        n = (int)std::thread::hardware_concurrency();
        // one process per GPU on a shared host (torchrun sets LOCAL_WORLD_SIZE): share the cores instead of oversubscribing them
        const char* lws = std::getenv("LOCAL_WORLD_SIZE");
        const int ranks_here = lws ? std::max(1, std::atoi(lws)) : 1;
        n = std::max(2, n / ranks_here);
        const char* e = std::getenv("QUILT_B200_HOST_THREADS");
        if (e) n = std::atoi(e);
        if (n < 1) n = 1;
        if (n > 16) n = 16;
    }
    return n;
}
// run f(i) for i in [0, n) on the host threads (staging copies are memory-bound: a few threads saturate DRAM)
template <class F>
void parallel_for(int n, F&& f) {
    const int nt = std::min(host_threads(), n);
    if (nt <= 1) {
        for (int i = 0; i < n; i++) f(i);
        return;
    }
    std::atomic<int> next{0};
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++)
        th.emplace_back([&]() {
            for (;;) {
                const int i = next.fetch_add(1);
                if (i >= n) break;
                f(i);
            }
        });
    for (auto& t : th) t.join();
}

// std::lgamma writes the global signgam: results are unpacked on several host threads, so use the re-entrant form
inline double lgamma_ts(double x) {
    int sign = 0;
    return ::lgamma_r(x, &sign);
}

// ------------------------------------------------------------------------------------------------ panel cache
// The device copy of a prepared reference is reused across calls.  It is identified by its CONTENT (shapes, ref_error
// and a 64-bit hash of every array), never by host addresses: a caller may rebuild the flat arrays per call (the Rcpp
// shim converts the rare/common lists into call-local vectors) or reuse freed addresses for a different panel.  At most
// PANEL_CACHE_MAX panels stay resident (least recently used is dropped).
constexpr size_t PANEL_CACHE_MAX = 4;
constexpr size_t PANEL_HASH_FULL = (size_t)256 << 20;  // arrays up to 256 MiB are hashed completely, larger ones in 4 KiB strides
struct PanelKey {
    int32_t K_full, nGrids, nSNPs, nMaxDH, n_special, nSNPs_all;
    double ref_error;
    uint64_t h[8];
    bool operator==(const PanelKey& o) const { return std::memcmp(this, &o, sizeof(PanelKey)) == 0; }
};
struct PanelEntry {
    PanelKey key;
    PanelDev dev;
    DBuf buf;
    DBuf hap_index;  // per-grid index of haplotypes sorted by symbol (built on the first full-panel pass)
    uint64_t last_use = 0;
};
std::vector<std::shared_ptr<PanelEntry>> g_panels;  // a staged batch holds a reference: eviction never frees a panel in use
uint64_t g_panel_clock = 0;

uint64_t hash_bytes(const void* p, size_t n) {
    if (!p || n == 0) return 0x9e3779b97f4a7c15ull;
    const unsigned char* b = (const unsigned char*)p;
    auto mix = [](uint64_t h, uint64_t v) {
        h ^= v;
        h *= 0xff51afd7ed558ccdull;
        h ^= h >> 32;
        return h;
    };
    uint64_t h = 0xcbf29ce484222325ull ^ (uint64_t)n;
    auto run = [&](size_t a, size_t e) {
        size_t i = a;
        for (; i + 8 <= e; i += 8) {
            uint64_t v;
            std::memcpy(&v, b + i, 8);
            h = mix(h, v);
        }
        for (; i < e; i++) h = mix(h, b[i]);
    };
    if (n <= PANEL_HASH_FULL) {
        run(0, n);
    } else {
        const size_t stride = 64 << 10;  // 4 KiB out of every 64 KiB, plus the tail
        for (size_t a = 0; a < n; a += stride) run(a, std::min(n, a + 4096));
        run(n - 4096, n);
    }
    return h;
}

// two argument structs of ONE batch describe the same host arrays (both are alive at the same time, so addresses are meaningful here)
bool same_panel_struct(const QuiltPanel& a, const QuiltPanel& b) {
    return a.K_full == b.K_full && a.nGrids == b.nGrids && a.nSNPs == b.nSNPs && a.nMaxDH == b.nMaxDH && a.hapMatcherR == b.hapMatcherR &&
           a.distinctHapsB == b.distinctHapsB && a.distinctHapsIE == b.distinctHapsIE && a.eMatDH_special_matrix == b.eMatDH_special_matrix &&
           a.n_special == b.n_special && a.eMatDH_special_matrix_helper == b.eMatDH_special_matrix_helper && a.ref_error == b.ref_error &&
           a.nSNPs_all == b.nSNPs_all && a.snp_is_common == b.snp_is_common && a.common_snp_index == b.common_snp_index &&
           a.rare_hap_offsets == b.rare_hap_offsets && a.rare_hap_snps == b.rare_hap_snps;
}

PanelKey panel_key(const QuiltPanel* p) {
    PanelKey k;
    std::memset(&k, 0, sizeof(k));
    k.K_full = p->K_full;
    k.nGrids = p->nGrids;
    k.nSNPs = p->nSNPs;
    k.nMaxDH = p->nMaxDH;
    k.n_special = p->n_special;
    k.nSNPs_all = p->nSNPs_all;
    k.ref_error = p->ref_error;
    const void* ptr[8] = {p->hapMatcherR, p->distinctHapsB, p->distinctHapsIE, p->eMatDH_special_matrix, p->eMatDH_special_matrix_helper,
                          p->snp_is_common, p->common_snp_index, p->rare_hap_offsets};
    const size_t nrare = (p->nSNPs_all > 0 && p->rare_hap_offsets) ? (size_t)std::max<int64_t>(p->rare_hap_offsets[p->K_full], 0) : 0;
    const size_t len[8] = {(size_t)p->K_full * p->nGrids, (size_t)p->nMaxDH * p->nGrids * 4, (size_t)p->nMaxDH * p->nSNPs * 8,
                           (size_t)std::max(p->n_special, 0) * 8, (size_t)p->nGrids * 8, (size_t)std::max(p->nSNPs_all, 0),
                           (size_t)std::max(p->nSNPs_all, 0) * 4, p->nSNPs_all > 0 ? (size_t)(p->K_full + 1) * 8 : 0};
    std::vector<std::thread> th;
    for (int i = 0; i < 8; i++) th.emplace_back([&, i]() { k.h[i] = hash_bytes(ptr[i], len[i]); });
    const uint64_t hr = hash_bytes(p->rare_hap_snps, nrare * 4);
    for (auto& t : th) t.join();
    k.h[7] ^= hr * 0x2545f4914f6cdd1dull;
    return k;
}

// The kernels derive a haplotype's emission at a SNP from its allele bit: (1 - eps) if set, eps otherwise.  The
// reference reads the same two numbers from distinctHapsIE (gibbs-small.cpp:208); check that the caller's matrix
// really is that function of distinctHapsB so the bit form is exact.
int check_distinctHapsIE(const QuiltPanel* p) {
    const double eps = p->ref_error, ome = 1 - p->ref_error;
    std::vector<int> used(p->nGrids, 0);
    for (int g = 0; g < p->nGrids; g++) {
        const uint8_t* col = p->hapMatcherR + (size_t)g * p->K_full;
        int m = 0;
        for (int k = 0; k < p->K_full; k++) m = std::max<int>(m, col[k]);
        if (m > p->nMaxDH) return set_err(QUILT_ERR_BAD_ARG, "hapMatcherR refers to a row beyond nMaxDH");
        used[g] = m;
    }
    for (int s = 0; s < p->nSNPs; s++) {
        const int g = s >> 5, b = s & 31;
        const double* ie = p->distinctHapsIE + (size_t)s * p->nMaxDH;
        const int32_t* wb = p->distinctHapsB + (size_t)g * p->nMaxDH;
        for (int i = 0; i < used[g]; i++) {
            const double want = (((uint32_t)wb[i] >> b) & 1u) ? ome : eps;
            if (ie[i] != want) return set_err(QUILT_ERR_UNSUPPORTED, "distinctHapsIE is not (bit ? 1 - ref_error : ref_error) of distinctHapsB");
        }
    }
    return QUILT_OK;
}

int get_panel(const QuiltPanel* p, PanelDev* out, std::shared_ptr<PanelEntry>* keep = nullptr) {
    if (p->K_full <= 0 || p->nGrids <= 0 || p->nSNPs <= 0 || p->nMaxDH <= 0 || p->nMaxDH > 255 || !p->hapMatcherR || !p->distinctHapsB ||
        !p->distinctHapsIE || !p->eMatDH_special_matrix_helper)
        return set_err(QUILT_ERR_BAD_ARG, "bad panel");
    if (p->nGrids != (p->nSNPs + 31) / 32) return set_err(QUILT_ERR_UNSUPPORTED, "panel grid must be 32 SNPs per grid");
    if (p->nSNPs_all > 0 && (!p->snp_is_common || !p->common_snp_index || !p->rare_hap_offsets)) return set_err(QUILT_ERR_BAD_ARG, "bad rare/common panel fields");
    const PanelKey key = panel_key(p);
    for (auto& e : g_panels)
        if (e->key == key) {
            e->last_use = ++g_panel_clock;
            *out = e->dev;
            if (keep) *keep = e;
            return QUILT_OK;
        }
    int rc = check_distinctHapsIE(p);
    if (rc != QUILT_OK) return rc;
    while (g_panels.size() >= PANEL_CACHE_MAX) {
        // drop the least recently used panel from the cache; its device memory goes when the last staged batch using it does
        size_t lru = 0;
        for (size_t i = 1; i < g_panels.size(); i++)
            if (g_panels[i]->last_use < g_panels[lru]->last_use) lru = i;
        CK(cudaStreamSynchronize(g_stream));
        g_panels.erase(g_panels.begin() + (long)lru);
    }
    auto e = std::make_shared<PanelEntry>();
    e->key = key;
    e->last_use = ++g_panel_clock;
    const size_t b_hm = al((size_t)p->K_full * p->nGrids), b_db = al((size_t)p->nMaxDH * p->nGrids * 4);
    const int nsp = std::max(p->n_special, 1);
    const size_t b_sp = al((size_t)nsp * 2 * 4), b_he = al((size_t)p->nGrids * 2 * 4), b_nu = al((size_t)p->nGrids * 4);
    std::vector<int32_t> n_used((size_t)p->nGrids, 0);
    for (int g = 0; g < p->nGrids; g++) {
        const uint8_t* col = p->hapMatcherR + (size_t)g * p->K_full;
        int m = 0;
        for (int k = 0; k < p->K_full; k++) m = std::max<int>(m, col[k]);
        n_used[(size_t)g] = std::min(m, p->nMaxDH);
    }
    size_t b_ic = 0, b_ci = 0, b_ro = 0, b_rs = 0, b_as = 0, b_ac = 0;
    int64_t n_rare = 0;
    std::vector<int8_t> asm_src;
    std::vector<int32_t> asm_cg0;
    if (p->nSNPs_all > 0) {
        if (!p->snp_is_common || !p->common_snp_index || !p->rare_hap_offsets) return set_err(QUILT_ERR_BAD_ARG, "bad rare/common panel fields");
        n_rare = p->rare_hap_offsets[p->K_full];
        b_ic = al(p->nSNPs_all);
        b_ci = al((size_t)p->nSNPs_all * 4);
        b_ro = al((size_t)(p->K_full + 1) * 8);
        b_rs = al((size_t)std::max<int64_t>(n_rare, 1) * 4);
        // where the SNPs of every all-SNP grid sit on the common axis (k_assemble_all)
        const int T_all = (p->nSNPs_all + 31) / 32;
        asm_src.assign((size_t)T_all * 32, (int8_t)-1);
        asm_cg0.assign((size_t)T_all, 0);
        for (int G = 0; G < T_all; G++) {
            int cg0 = -1;
            for (int b = 0; b < 32; b++) {
                const int s = 32 * G + b;
                if (s >= p->nSNPs_all || !p->snp_is_common[s]) continue;
                const int cs = p->common_snp_index[s] - 1;
                if (cs < 0 || cs >= p->nSNPs) return set_err(QUILT_ERR_BAD_ARG, "common_snp_index out of range");
                if (cg0 < 0) cg0 = cs >> 5;
                const int wsel = (cs >> 5) - cg0;
                if (wsel < 0 || wsel > 1) return set_err(QUILT_ERR_BAD_ARG, "common_snp_index is not increasing along the all-SNP axis");
                asm_src[(size_t)G * 32 + b] = (int8_t)((wsel << 5) | (cs & 31));
            }
            asm_cg0[(size_t)G] = std::max(cg0, 0);
        }
        b_as = al(asm_src.size());
        b_ac = al(asm_cg0.size() * 4);
    }
    CK(e->buf.alloc(b_hm + b_db + b_sp + b_he + b_nu + b_ic + b_ci + b_ro + b_rs + b_as + b_ac));
    char* d = (char*)e->buf.p;
    PanelDev& D = e->dev;
    D.K_full = p->K_full;
    D.Tc = p->nGrids;
    D.nSNPsC = p->nSNPs;
    D.nMaxDH = p->nMaxDH;
    D.n_special = nsp;
    D.nSNPs_all = p->nSNPs_all;
    D.ref_error = p->ref_error;
    D.hapMatcherR = (const uint8_t*)d;
    CK(cudaMemcpy(d, p->hapMatcherR, (size_t)p->K_full * p->nGrids, cudaMemcpyHostToDevice));
    d += b_hm;
    D.distinctHapsB = (const int32_t*)d;
    CK(cudaMemcpy(d, p->distinctHapsB, (size_t)p->nMaxDH * p->nGrids * 4, cudaMemcpyHostToDevice));
    d += b_db;
    D.special = (const int32_t*)d;
    if (p->n_special > 0)
        CK(cudaMemcpy(d, p->eMatDH_special_matrix, (size_t)p->n_special * 2 * 4, cudaMemcpyHostToDevice));
    else
        CK(cudaMemset(d, 0, b_sp));
    d += b_sp;
    D.helper = (const int32_t*)d;
    CK(cudaMemcpy(d, p->eMatDH_special_matrix_helper, (size_t)p->nGrids * 2 * 4, cudaMemcpyHostToDevice));
    d += b_he;
    D.n_used = (const int32_t*)d;
    CK(cudaMemcpy(d, n_used.data(), (size_t)p->nGrids * 4, cudaMemcpyHostToDevice));
    d += b_nu;
    D.asm_src = nullptr;
    D.asm_cg0 = nullptr;
    D.snp_is_common = nullptr;
    D.common_snp_index = nullptr;
    D.rare_off = nullptr;
    D.rare_snps = nullptr;
    if (p->nSNPs_all > 0) {
        D.snp_is_common = (const uint8_t*)d;
        CK(cudaMemcpy(d, p->snp_is_common, p->nSNPs_all, cudaMemcpyHostToDevice));
        d += b_ic;
        D.common_snp_index = (const int32_t*)d;
        CK(cudaMemcpy(d, p->common_snp_index, (size_t)p->nSNPs_all * 4, cudaMemcpyHostToDevice));
        d += b_ci;
        D.rare_off = (const int64_t*)d;
        CK(cudaMemcpy(d, p->rare_hap_offsets, (size_t)(p->K_full + 1) * 8, cudaMemcpyHostToDevice));
        d += b_ro;
        D.rare_snps = (const int32_t*)d;
        if (n_rare > 0) CK(cudaMemcpy(d, p->rare_hap_snps, (size_t)n_rare * 4, cudaMemcpyHostToDevice));
        d += b_rs;
        D.asm_src = (const int8_t*)d;
        CK(cudaMemcpy(d, asm_src.data(), asm_src.size(), cudaMemcpyHostToDevice));
        d += b_as;
        D.asm_cg0 = (const int32_t*)d;
        CK(cudaMemcpy(d, asm_cg0.data(), asm_cg0.size() * 4, cudaMemcpyHostToDevice));
        d += b_ac;
    }
    *out = D;
    if (keep) *keep = e;
    g_panels.push_back(std::move(e));
    return QUILT_OK;
}

// ------------------------------------------------------------------------------------------------ geometry
struct Geo {
    int NT, EPT;
    int CL = 1;  // CTAs per job in the sweep / shard kernels (thread-block cluster splitting the K states)
};
bool pick_geo(int K, Geo* g, int NH = 2) {
    if (NH == 3 && K > 2048) return false;  // three eMatGrid columns per stage: the shared-memory ring holds K <= 2048
    // QUILT_B200_GEO=NTxEPT overrides (experiments); default: few fat warps — the per-read chain is latency-bound, fewer
    // warps mean less redundant scalar work and a shorter cross-warp reduction, more elements per thread mean more ILP
    static int env_nt = -1, env_ept = 0;
    if (env_nt < 0) {
        env_nt = 0;
        const char* e = std::getenv("QUILT_B200_GEO");
        if (e) std::sscanf(e, "%dx%d", &env_nt, &env_ept);
    }
    if (env_nt > 0 && env_nt * env_ept >= K) {
        *g = {env_nt, env_ept, 1};
        return true;
    }
    // eight elements per thread below K = 4096 (measured on B200 against 4 and 16 per thread: 7 % / 14 % / 21 % faster
    // than the 4-per-thread geometries at K = 512 / 1024 / 2048, 16 per thread slower at every K <= 2048)
    if (K <= 512)
        *g = {64, 8};
    else if (K <= 1024)
        *g = {128, 8};
    else if (K <= 2048)
        *g = {256, 8};
    else if (K <= 4096)
        *g = {256, 16};  // 512 x 8 measured 39 % slower on the K = 4096 benchmark (128-register cap)
    else if (K <= 8192 && NH == 2)
        *g = {256, 16, 2};  // two-CTA cluster, 4096 states per CTA
    else
        return false;
    return true;
}

template <typename F>
int with_geo(const Geo& g, F&& f) {
#define QB_GEO(NT_, EPT_) \
    if (g.NT == NT_ && g.EPT == EPT_) return f(std::integral_constant<int, NT_>(), std::integral_constant<int, EPT_>());
    QB_GEO(64, 8)
    QB_GEO(128, 8)
    QB_GEO(256, 8)
    QB_GEO(256, 16)
    QB_GEO(512, 8)  // experiments only (QUILT_B200_GEO=512x8): the comparison quoted in DESIGN.md
#undef QB_GEO
    return set_err(QUILT_ERR_UNSUPPORTED, "no kernel geometry");
}

// ------------------------------------------------------------------------------------------------ jobs
struct JobLayoutIn {  // byte offsets inside the job's input region
    size_t which, rs, roff, u, pRA, wif0, ts, ginfo, dense_reads, runif_reads, runif_shard, tm, desc, H0, runif_block, runif_H_class, L_grid, end;
};
// Results live in two device zones: zone A is copied to the host, zone B (endB bytes per job) never leaves the device.
// genProbsM_t / genProbsF_t go to zone B whenever there is ONE sampling sweep — they are then an exact function of
// hapProbs_t (gibbs-small.cpp:621-633) and the host re-forms them when unpacking (two thirds of the bulk D2H saved);
// with QUILT_F_OUTPUT_NO_PROBS hapProbs_t stays in zone B as well (intermediate call of a device-resident chain).
struct JobLayoutOut {
    size_t underflow, lik, hap, genM, genF, H, Hclass, cat, Hs, end;
    size_t endB = 0;
    bool hap_dev_only = false, gen_dev_only = false;
};

struct HostJob {
    QuiltGibbsArgs a;
    int R = 0, nU = 0, n_dense = 0, n_tab = 0, n_ep = 0, n_its = 0;
    int bucket = -1;
    JobLayoutIn li;
    JobLayoutOut lo;
    size_t in_off = 0, out_off = 0, outB_off = 0;  // offsets of the job's regions in the batch arenas
    std::vector<ReadDesc> desc;
    std::vector<int32_t> rs, ts, ginfo, dense_reads;
    // debug copies (QUILT_F_RETURN_ALPHA / _EXTRA)
    std::vector<double> dbg_alpha, dbg_beta, dbg_eG, dbg_c, dbg_eMatRead;
};

struct Bucket {
    BatchParams P;
    Geo geo;
    std::vector<int> jobs;
    std::vector<int> block_its;
    int n_slots = 0;
    int R_max = 0, n_tab_max = 0, n_dense_max = 0;
    size_t slot_bytes = 0;
    // slot layout
    size_t o_alpha, o_beta, o_eG, o_c, o_W, o_Wc, o_tabs, o_dense, o_xprob, o_snp_type, o_rate, o_hapLocal, o_blk;
    size_t o_cinfo, o_cperm, o_ccls, o_cent, o_crec;  // haplotype classes (classes.cuh)
    bool classes = false;
    char* slots = nullptr;  // into the batch's slot arena (buckets run one after the other and share it)
    DBuf djobs;
    HBuf hjobs;
    bool perform_block = false, do_shard = false, debug = false;
    ~Bucket() {
        cached_release(djobs, g_dcache_jobs);
        cached_release(hjobs, g_hcache_jobs);
    }
};

}  // namespace

struct QuiltGpuBatch {
    int n = 0;
    std::vector<HostJob> jobs;
    std::vector<std::unique_ptr<Bucket>> buckets;
    PanelDev panel;
    std::shared_ptr<PanelEntry> panel_ref;
    std::unique_ptr<DBuf> din_, dout_;
    size_t slot_arena_bytes = 0;  // what this batch needs of the shared wave-state arena (g_slots)
    DBuf doutB;  // zone B of the results (device only)
    size_t outB_bytes = 0;
    std::unique_ptr<HBuf> hin_, hout_;
    DBuf& din() const { return *din_; }
    DBuf& dout() const { return *dout_; }
    HBuf& hin() const { return *hin_; }
    HBuf& hout() const { return *hout_; }
    size_t in_bytes = 0, out_bytes = 0;
    bool ran = false, fetched_raw = false;
    double chain_ms = 0;   // device time of the selection that produced this batch's haplotype lists
    bool chained = false;  // which_haps_to_use of the jobs were written on the device (quilt_gpu_batch_chain_select)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> sweep_events;
    double total_ms = 0, sweep_ms = 0;
    int n_sweep_launches = 0;
    ~QuiltGpuBatch() {
        cached_release(doutB, g_dcache_outB);
        g_dcache_in.release(std::move(din_));
        g_dcache_out.release(std::move(dout_));
        g_hcache_in.release(std::move(hin_));
        g_hcache_out.release(std::move(hout_));
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        for (auto& p : sweep_events) {
            cudaEventDestroy(p.first);
            cudaEventDestroy(p.second);
        }
    }
};

namespace {

// reads with up to this many SNPs keep a 2^nb-entry emission table, longer ones a dense column (QUILT_B200_NBMAX: experiments)
int nb_limit() {
    static int v = -1;
    if (v < 0) {
        v = NBMAX;
        const char* e = std::getenv("QUILT_B200_NBMAX");
        if (e) v = std::max(1, std::min(NBMAX, std::atoi(e)));
    }
    return v;
}

// classify every read and lay out its emission table; host-side, O(sum J)
int prepare_job(HostJob& j) {
    const QuiltGibbsArgs& a = j.a;
    const int R = a.reads.nReads, T = a.nGrids;
    j.R = R;
    j.n_its = a.n_gibbs_burn_in_its + a.n_gibbs_sample_its;
    if (R <= 0) return set_err(QUILT_ERR_BAD_ARG, "no reads");
    if (!a.reads.offsets || !a.reads.u || !a.reads.bq || !a.reads.wif0 || !a.which_haps_to_use || !a.transMatRate_tc_H || !a.H0)
        return set_err(QUILT_ERR_BAD_ARG, "null input array");
    if (j.n_its > 0 && !a.runif_reads) return set_err(QUILT_ERR_BAD_ARG, "runif_reads missing");
    j.nU = a.reads.offsets[R];
    j.rs.assign(T + 1, 0);
    int prev = 0;
    for (int r = 0; r < R; r++) {
        const int w = a.reads.wif0[r];
        if (w < prev || w >= T || w < 0) return set_err(QUILT_ERR_BAD_ARG, "wif0 must be non-decreasing and inside [0, nGrids)");
        prev = w;
        j.rs[w + 1]++;
    }
    for (int g = 0; g < T; g++) j.rs[g + 1] += j.rs[g];
    j.desc.assign(R, ReadDesc());
    j.ts.assign(T + 1, 0);
    j.dense_reads.clear();
    int n_tab = 0;
    int g_cur = 0;
    for (int r = 0; r < R; r++) {
        const int w = a.reads.wif0[r];
        while (g_cur < w) j.ts[++g_cur] = n_tab;
        ReadDesc& d = j.desc[r];
        std::memset(&d, 0, sizeof(d));
        const int o = a.reads.offsets[r];
        int cnt = a.reads.offsets[r + 1] - o;
        if (cnt <= 0) return set_err(QUILT_ERR_BAD_ARG, "read without SNPs");
        if (cnt - 1 >= a.Jmax) cnt = a.Jmax + 1;
        bool table = cnt <= nb_limit();
        bool run = true;
        for (int q = 0; q < cnt && table; q++) {
            const int s = a.reads.u[o + q];
            if (s < 0 || s >= a.nSNPs) return set_err(QUILT_ERR_BAD_ARG, "SNP index outside [0, nSNPs)");
            const int wg = s >> 5;
            if (wg < w - 1 || wg > w + 1) table = false;
            if (q > 0 && s != a.reads.u[o + q - 1] + 1) run = false;
        }
        // non-consecutive SNPs (possible only for hand-made inputs: a read's SNPs are a contiguous index range) would
        // need a per-haplotype gathered pattern; such reads are kept as dense columns instead
        if (table && !run) table = false;
        if (table) {
            d.nb = (uint8_t)cnt;
            d.off = (uint32_t)n_tab;
            n_tab += 1 << cnt;
            d.tnext = (uint32_t)n_tab;
            const int s0 = a.reads.u[o];
            if (run) {
                d.mode = MODE_RUN;
                d.g0rel = (int8_t)((s0 >> 5) - w);
                d.b0 = (uint8_t)(s0 & 31);
            } else {
                d.mode = MODE_GATHER;
                for (int q = 0; q < cnt; q++) {
                    const int s = a.reads.u[o + q];
                    d.sel[q] = (uint8_t)((((s >> 5) - w + 1) << 5) | (s & 31));
                }
            }
        } else {
            for (int q = 0; q < cnt; q++) {
                const int s = a.reads.u[o + q];
                if (s < 0 || s >= a.nSNPs) return set_err(QUILT_ERR_BAD_ARG, "SNP index outside [0, nSNPs)");
            }
            d.mode = MODE_DENSE;
            d.tnext = (uint32_t)n_tab;
            d.off = (uint32_t)j.dense_reads.size();
            j.dense_reads.push_back(r);
        }
    }
    while (g_cur < T) j.ts[++g_cur] = n_tab;
    j.n_tab = n_tab;
    j.n_dense = (int)j.dense_reads.size();
    // staging chunks of the sweep kernel: greedy packing of every grid's reads into SW_MAXR reads / SW_MAXTAB entries
    j.ginfo.assign((size_t)(T + 1) * 4, 0);
    for (int g = 0; g <= T; g++) {
        j.ginfo[4 * g + 0] = j.rs[g];
        j.ginfo[4 * g + 1] = j.ts[g];
        j.ginfo[4 * g + 3] = j.ts[g];
    }
    for (int g = 0; g < T; g++) {
        const int r0 = j.rs[g], r1 = j.rs[g + 1];
        int c0 = 0;
        uint32_t tab0 = (uint32_t)j.ts[g];
        while (r0 + c0 < r1) {
            int n = 0;
            uint32_t tend = tab0;
            while (r0 + c0 + n < r1 && n < SW_MAXR) {
                const uint32_t tn = j.desc[r0 + c0 + n].tnext;
                if (tn - tab0 > (uint32_t)SW_MAXTAB) break;
                tend = tn;
                n++;
            }
            if (n == 0) return set_err(QUILT_ERR_UNSUPPORTED, "emission table larger than the staging buffer");
            j.desc[r0 + c0].chunk_n = (uint16_t)n;
            j.desc[r0 + c0].chunk_tend = tend;
            if (c0 == 0) {
                j.ginfo[4 * g + 2] = n;
                j.ginfo[4 * g + 3] = (int32_t)tend;
            }
            tab0 = tend;
            c0 += n;
        }
    }
    return QUILT_OK;
}

// upper bound of the episode-stream values one call can consume (quilt_b200.h, QuiltGibbsArgs.unif_stream)
size_t episode_stream_need(const QuiltGibbsArgs& a) {
    const bool diploid = (a.flags & QUILT_F_SAMPLE_IS_DIPLOID) != 0;
    const bool shard = (a.flags & QUILT_F_DO_SHARD_BLOCK_GIBBS) != 0;
    if (!(a.flags & QUILT_F_PERFORM_BLOCK_GIBBS)) return 0;
    const size_t R = (size_t)a.reads.nReads;
    return (size_t)a.n_block_gibbs_iterations * (8 * R + (diploid ? 0 : R) + (shard ? (size_t)std::max(a.nGrids - 1, 0) : 0));
}
// stream position of episode e for a DIPLOID call (static: no data-dependent draws)
size_t diploid_episode_base(const QuiltGibbsArgs& a, int e) {
    const bool shard = (a.flags & QUILT_F_DO_SHARD_BLOCK_GIBBS) != 0;
    return (size_t)e * (8 * (size_t)a.reads.nReads + (shard ? (size_t)std::max(a.nGrids - 1, 0) : 0));
}

void layout_in(HostJob& j) {
    const QuiltGibbsArgs& a = j.a;
    const int R = j.R, T = a.nGrids;
    j.n_ep = a.n_block_gibbs_iterations;
    size_t o = 0;
    JobLayoutIn& L = j.li;
    L.which = o, o += al((size_t)a.K * 4);
    L.rs = o, o += al((size_t)(T + 1) * 4);
    L.roff = o, o += al((size_t)(R + 1) * 4);
    L.u = o, o += al((size_t)j.nU * 4);
    L.pRA = o, o += al((size_t)j.nU * 16);
    L.wif0 = o, o += al((size_t)R * 4);
    L.ts = o, o += al((size_t)(T + 1) * 4);
    L.ginfo = o, o += al((size_t)(T + 1) * 16);
    L.dense_reads = o, o += al((size_t)std::max(j.n_dense, 1) * 4);
    L.runif_reads = o, o += al((size_t)std::max(j.n_its, 1) * R * 8);
    L.runif_shard = o, o += al((size_t)std::max(j.n_ep, 1) * std::max(T - 1, 1) * 8);
    L.tm = o, o += al((size_t)std::max(T - 1, 1) * 16);
    L.desc = o, o += al((size_t)R * sizeof(ReadDesc));
    L.H0 = o, o += al((size_t)R * 4);
    const bool nipt_block = !(a.flags & QUILT_F_SAMPLE_IS_DIPLOID) && (a.flags & QUILT_F_PERFORM_BLOCK_GIBBS) && j.n_ep > 0;
    // episode-stream mode (quilt_b200.h): the three-haplotype kernels walk the caller's flat stream themselves (the
    // number of H_class draws per episode is data-dependent); it travels in the runif_block slot
    const size_t nipt_stream = (nipt_block && a.unif_stream) ? episode_stream_need(a) : 0;
    L.runif_block = o, o += al(nipt_block ? (nipt_stream ? nipt_stream * 8 : (size_t)j.n_ep * R * 8) : 0);
    L.runif_H_class = o, o += al((nipt_block && !nipt_stream) ? (size_t)j.n_ep * R * 8 : 0);
    L.L_grid = o, o += al(nipt_block ? (size_t)T * 4 : 0);
    L.end = o;
    JobLayoutOut& O = j.lo;
    o = 0;
    O.underflow = o, o += 256;
    O.lik = o, o += al((size_t)std::max(j.n_its, 1) * LIK_N * 8);
    O.hap_dev_only = (a.flags & QUILT_F_OUTPUT_NO_PROBS) != 0;
    O.gen_dev_only = O.hap_dev_only || a.n_gibbs_sample_its == 1;
    size_t ob = 0;
    if (O.hap_dev_only)
        O.hap = ob, ob += al((size_t)a.nSNPs * 24);
    else
        O.hap = o, o += al((size_t)a.nSNPs * 24);
    if (O.gen_dev_only) {
        O.genM = ob, ob += al((size_t)a.nSNPs * 24);
        O.genF = ob, ob += al((size_t)a.nSNPs * 24);
    } else {
        O.genM = o, o += al((size_t)a.nSNPs * 24);
        O.genF = o, o += al((size_t)a.nSNPs * 24);
    }
    O.endB = ob;
    O.H = o, o += al((size_t)R * 4);
    O.Hclass = o, o += al((size_t)R * 4);
    O.cat = o, o += al((size_t)R * 4);
    // labels after each sampling sweep (double_list_of_ending_read_labels); with one sampling sweep they are H itself
    O.Hs = o, o += al(a.n_gibbs_sample_its > 1 ? (size_t)a.n_gibbs_sample_its * R * 4 : 0);
    O.end = o;
}

// (pR, pA) per read-SNP exactly as the reference walks them (gibbs-small.cpp:172-181): the pair is only updated
// for bq != 0 and is carried over from the previous SNP / previous read otherwise.
void fill_pRA(const HostJob& j, double* pRA) {
    const QuiltGibbsArgs& a = j.a;
    double eps, pR = 1, pA = 1;
    for (int r = 0; r < j.R; r++) {
        const int o = a.reads.offsets[r];
        int cnt = a.reads.offsets[r + 1] - o;
        const int used = std::min(cnt, a.Jmax + 1);
        for (int q = 0; q < cnt; q++) {
            if (q < used) {
                const int bq = a.reads.bq[o + q];
                if (bq < 0) {
                    eps = std::pow(10, (double(bq) / 10));
                    pR = 1 - eps;
                    pA = eps / 3;
                }
                if (bq > 0) {
                    eps = std::pow(10, (-double(bq) / 10));
                    pR = eps / 3;
                    pA = 1 - eps;
                }
            }
            pRA[2 * (size_t)(o + q)] = pR;
            pRA[2 * (size_t)(o + q) + 1] = pA;
        }
    }
}

void fill_in(const HostJob& j, char* base) {
    const QuiltGibbsArgs& a = j.a;
    const JobLayoutIn& L = j.li;
    const int R = j.R, T = a.nGrids;
    std::memcpy(base + L.which, a.which_haps_to_use, (size_t)a.K * 4);
    std::memcpy(base + L.rs, j.rs.data(), (size_t)(T + 1) * 4);
    std::memcpy(base + L.roff, a.reads.offsets, (size_t)(R + 1) * 4);
    std::memcpy(base + L.u, a.reads.u, (size_t)j.nU * 4);
    fill_pRA(j, reinterpret_cast<double*>(base + L.pRA));
    std::memcpy(base + L.wif0, a.reads.wif0, (size_t)R * 4);
    std::memcpy(base + L.ts, j.ts.data(), (size_t)(T + 1) * 4);
    std::memcpy(base + L.ginfo, j.ginfo.data(), (size_t)(T + 1) * 16);
    if (j.n_dense) std::memcpy(base + L.dense_reads, j.dense_reads.data(), (size_t)j.n_dense * 4);
    if (j.n_its > 0) std::memcpy(base + L.runif_reads, a.runif_reads, (size_t)j.n_its * R * 8);
    if (j.n_ep > 0 && T > 1 && a.unif_stream && (a.flags & QUILT_F_SAMPLE_IS_DIPLOID) && (a.flags & QUILT_F_DO_SHARD_BLOCK_GIBBS)) {
        // episode stream, diploid: the shard uniforms of episode e follow its 8 * nReads block-Gibbs uniforms
        for (int e = 0; e < j.n_ep; e++)
            std::memcpy(base + L.runif_shard + (size_t)e * (T - 1) * 8, a.unif_stream + diploid_episode_base(a, e) + 8 * (size_t)R, (size_t)(T - 1) * 8);
    } else if (j.n_ep > 0 && T > 1 && a.runif_shard)
        std::memcpy(base + L.runif_shard, a.runif_shard, (size_t)j.n_ep * (T - 1) * 8);
    if (T > 1) std::memcpy(base + L.tm, a.transMatRate_tc_H, (size_t)(T - 1) * 16);
    std::memcpy(base + L.desc, j.desc.data(), (size_t)R * sizeof(ReadDesc));
    std::memcpy(base + L.H0, a.H0, (size_t)R * 4);
    const bool nipt_block = !(a.flags & QUILT_F_SAMPLE_IS_DIPLOID) && (a.flags & QUILT_F_PERFORM_BLOCK_GIBBS) && j.n_ep > 0;
    if (nipt_block) {
        if (a.unif_stream) {
            std::memcpy(base + L.runif_block, a.unif_stream, episode_stream_need(a) * 8);
        } else {
            std::memcpy(base + L.runif_block, a.runif_block, (size_t)j.n_ep * R * 8);
            std::memcpy(base + L.runif_H_class, a.runif_H_class, (size_t)j.n_ep * R * 8);
        }
        std::memcpy(base + L.L_grid, a.L_grid, (size_t)T * 4);
    }
}

typedef std::tuple<int, int, int, int, uint32_t, int, int, double, double, double, double, int, std::vector<int>> BucketKey;

BucketKey bucket_key(const QuiltGibbsArgs& a) {
    std::vector<int> bi(a.block_gibbs_iterations, a.block_gibbs_iterations + a.n_block_gibbs_iterations);
    return BucketKey(a.K, a.nGrids, a.nSNPs, a.shuffle_bin_radius, a.flags & ~(uint32_t)QUILT_F_OUTPUT_NO_PROBS, a.n_gibbs_burn_in_its, a.n_gibbs_sample_its, a.ff,
                     a.maxDifferenceBetweenReads, a.class_sum_cutoff, a.block_gibbs_quantile_prob, a.Jmax, bi);
}

int validate(const QuiltGibbsArgs& a) {
    if (!a.panel) return set_err(QUILT_ERR_BAD_ARG, "panel is NULL");
    if (a.K <= 0 || a.nGrids <= 0 || a.nSNPs <= 0) return set_err(QUILT_ERR_BAD_ARG, "bad K / nGrids / nSNPs");
    if (!a.which_haps_to_use || (!a.transMatRate_tc_H && a.nGrids > 1)) return set_err(QUILT_ERR_BAD_ARG, "which_haps_to_use / transMatRate_tc_H is NULL");
    if (a.reads.nReads < 0 || (a.reads.nReads > 0 && (!a.reads.offsets || !a.reads.u || !a.reads.bq || !a.reads.wif0 || !a.H0)))
        return set_err(QUILT_ERR_BAD_ARG, "reads / H0 arrays are NULL");
    if (a.n_gibbs_burn_in_its + a.n_gibbs_sample_its > 0 && a.reads.nReads > 0 && !a.runif_reads) return set_err(QUILT_ERR_BAD_ARG, "runif_reads is NULL");
    if (a.unif_stream && (size_t)std::max<int64_t>(a.n_unif_stream, 0) < episode_stream_need(a))
        return set_err(QUILT_ERR_BAD_ARG, "unif_stream is shorter than n_episodes * (8 nReads [+ nReads] [+ nGrids - 1])");
    if (a.nGrids != (a.nSNPs + 31) / 32) return set_err(QUILT_ERR_UNSUPPORTED, "grid must be 32 SNPs per grid (grid32)");
    if (a.K > 8192) return set_err(QUILT_ERR_UNSUPPORTED, "Ksubset > 8192 not supported");
    const bool diploid = (a.flags & QUILT_F_SAMPLE_IS_DIPLOID) != 0;
    if (diploid && a.ff != 0) return set_err(QUILT_ERR_BAD_ARG, "sample_is_diploid with ff != 0");
    if (!diploid) {
        if (a.ff < 0 || a.ff >= 1) return set_err(QUILT_ERR_BAD_ARG, "ff outside [0, 1)");
        if (a.K > 2048) return set_err(QUILT_ERR_UNSUPPORTED, "three-haplotype (NIPT) calls support Ksubset <= 2048");
        if ((a.flags & QUILT_F_PERFORM_BLOCK_GIBBS) && a.n_block_gibbs_iterations > 0) {
            if ((!a.unif_stream && (!a.runif_block || !a.runif_H_class)) || !a.L_grid)
                return set_err(QUILT_ERR_BAD_ARG, "NIPT block Gibbs needs L_grid and either unif_stream or runif_block + runif_H_class");
        }
        if (a.flags & QUILT_F_DO_SHARD_BLOCK_GIBBS) return set_err(QUILT_ERR_UNSUPPORTED, "shard pass is diploid-only (functions.R:2552-2556)");
    }
    const bool rc = (a.flags & QUILT_F_MAKE_EMATREAD_RARE_COMMON) != 0;
    if (rc) {
        if (a.panel->nSNPs_all != a.nSNPs) return set_err(QUILT_ERR_BAD_ARG, "rare/common call: nSNPs must equal panel nSNPs_all");
    } else {
        if (a.panel->nSNPs != a.nSNPs) return set_err(QUILT_ERR_BAD_ARG, "common-SNP call: nSNPs must equal panel nSNPs");
    }
    if ((a.flags & QUILT_F_PERFORM_BLOCK_GIBBS) && a.n_block_gibbs_iterations > 0) {
        if ((a.flags & QUILT_F_DO_SHARD_BLOCK_GIBBS) && !(a.flags & QUILT_F_SHARD_CHECK_EVERY_PAIR))
            return set_err(QUILT_ERR_UNSUPPORTED, "shard pass without shard_check_every_pair not supported yet");
        if ((a.flags & QUILT_F_DO_SHARD_BLOCK_GIBBS) && !a.runif_shard && !a.unif_stream) return set_err(QUILT_ERR_BAD_ARG, "runif_shard missing");
    }
    if (a.Jmax < 0) return set_err(QUILT_ERR_BAD_ARG, "Jmax < 0");
    for (int k = 0; k < a.K; k++)
        if (a.which_haps_to_use[k] < 1 || a.which_haps_to_use[k] > a.panel->K_full) return set_err(QUILT_ERR_BAD_ARG, "which_haps_to_use out of range");
    return QUILT_OK;
}

void make_params(const QuiltGibbsArgs& a, BatchParams* P) {
    std::memset(P, 0, sizeof(*P));
    P->K = a.K;
    P->Kp = (a.K + 31) & ~31;
    P->T = a.nGrids;
    P->NH = (a.flags & QUILT_F_SAMPLE_IS_DIPLOID) ? 2 : 3;
    P->nSNPs = a.nSNPs;
    P->n_its = a.n_gibbs_burn_in_its + a.n_gibbs_sample_its;
    P->n_burn = a.n_gibbs_burn_in_its;
    P->flags = a.flags;
    P->ff = a.ff;
    P->one_over_K = 1 / double(a.K);
    P->d2 = 1 / a.maxDifferenceBetweenReads;
    P->class_sum_cutoff = a.class_sum_cutoff;
    const double pp[3] = {0.5, (1 - a.ff) / 2, (a.ff / 2)};
    for (int i = 0; i < 3; i++) P->prior[i] = pp[i];
    const double rlc[7][3] = {{1, 0, 0},
                              {0, 1, 0},
                              {0, 0, 1},
                              {pp[0] / (pp[0] + pp[1]), pp[1] / (pp[0] + pp[1]), 0},
                              {pp[0] / (pp[0] + pp[2]), 0, pp[2] / (pp[0] + pp[2])},
                              {0, pp[1] / (pp[1] + pp[2]), pp[2] / (pp[1] + pp[2])},
                              {pp[0], pp[1], pp[2]}};
    std::memcpy(P->rlc, rlc, sizeof(rlc));
    P->ref_error = a.panel->ref_error;
    P->rare_common = (a.flags & QUILT_F_MAKE_EMATREAD_RARE_COMMON) ? 1 : 0;
    P->Jmax = a.Jmax;
    P->shuffle_bin_radius = a.shuffle_bin_radius;
    P->block_q = a.block_gibbs_quantile_prob;
    {
        // rcpp_get_log_p_H_class2 (gibbs-nipt-block.cpp:169-208): terms of n1 .. n6, branch chosen by ff
        const double ff = a.ff;
        P->lhc[0] = std::log(0.5);
        P->lhc[1] = (ff == 1) ? std::log(0.001) : std::log(0.5 - ff * 0.5);
        P->lhc[2] = (ff == 0) ? std::log(0.001) : std::log(ff * 0.5);
        P->lhc[3] = std::log(1 - ff * 0.5);
        P->lhc[4] = std::log(1 * 0.5 + ff * 0.5);
        P->lhc[5] = std::log(1 * 0.5);
    }
    const char* bm = std::getenv("QUILT_B200_DBG");
    P->dbg = bm ? (uint32_t)std::atoi(bm) : 0u;
    // grids with fewer visited reads walk all K states: the class totals cost about as much as six K-long reads (QUILT_B200_CLS_MIN: experiments)
    const char* cm = std::getenv("QUILT_B200_CLS_MIN");
    P->cls_min_reads = cm ? std::atoi(cm) : 8;
}

// the three-haplotype (NIPT) instance exists for geometries whose shared-memory ring fits (K <= 2048)
template <int NT, int EPT>
constexpr bool nipt_geo() {
    return NT * EPT <= 2048;
}
// launch of a kernel whose CTAs come in clusters of CL (CL = 1: plain launch)
template <typename... KArgs, typename... Args>
cudaError_t launch_cl(void (*kern)(KArgs...), int n_ctas, int nt, size_t smem, int CL, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_ctas);
    cfg.blockDim = dim3(nt);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = g_stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}

template <int NT, int EPT>
int sweep_occupancy(int Kp, int NH, int CL, bool cls, int* occ, int* smem) {
    const SweepSmemLayout L = sweep_smem_layout(NT * EPT, NH, NT, CL);
    *smem = L.total;
    if (CL == 2) {
        if constexpr (NT == 256 && EPT == 16) {
            if (NH != 2) return set_err(QUILT_ERR_UNSUPPORTED, "cluster sweep is diploid-only");
            CK(cudaFuncSetAttribute(k_sweep<NT, EPT, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
            *occ = 1;  // one CTA per SM; a job takes two SMs
            return QUILT_OK;
        } else {
            return set_err(QUILT_ERR_UNSUPPORTED, "no cluster sweep kernel for this geometry");
        }
    }
    if (NH == 2 && cls) {
        CK(cudaFuncSetAttribute(k_sweep<NT, EPT, 2, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k_sweep<NT, EPT, 2, 1, true>, NT, L.total));
    } else if (NH == 2) {
        CK(cudaFuncSetAttribute(k_sweep<NT, EPT, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k_sweep<NT, EPT, 2>, NT, L.total));
    } else if constexpr (nipt_geo<NT, EPT>()) {
        CK(cudaFuncSetAttribute(k_sweep<NT, EPT, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k_sweep<NT, EPT, 3>, NT, L.total));
    } else {
        return set_err(QUILT_ERR_UNSUPPORTED, "no three-haplotype sweep kernel for this K");
    }
    return QUILT_OK;
}

int setup_bucket(QuiltGpuBatch* B, Bucket& bk, size_t* mem_budget) {
    const BatchParams& P = bk.P;
    if (!pick_geo(P.K, &bk.geo, P.NH)) return set_err(QUILT_ERR_UNSUPPORTED, "Ksubset too large");
    for (int ji : bk.jobs) bk.R_max = std::max(bk.R_max, B->jobs[ji].R);
    // haplotype classes in the sweep kernel (classes.cuh): diploid calls, one CTA per job, and enough reads per grid for the
    // per-grid class totals to pay (common-SNP calls: ~16 visited reads per grid; all-SNP calls: ~6, they keep the K-long instance)
    {
        const char* e = std::getenv("QUILT_B200_CLASSES");  // tests / experiments: 0 = never, 1 = whenever possible
        const int cls_env = e ? 1 + std::atoi(e) : 0;
        bk.classes = (P.NH == 2 && bk.geo.CL == 1) && (cls_env == 2 || (cls_env == 0 && (double)bk.R_max >= 10.0 * P.T));
    }
    int occ = 0, smem = 0;
    int rc = with_geo(bk.geo, [&](auto nt, auto ept) { return sweep_occupancy<decltype(nt)::value, decltype(ept)::value>(P.Kp, P.NH, bk.geo.CL, bk.classes, &occ, &smem); });
    if (rc != QUILT_OK) return rc;
    if (occ < 1) return set_err(QUILT_ERR_UNSUPPORTED, "sweep kernel does not fit on an SM for this K");
    for (int ji : bk.jobs) {
        const HostJob& j = B->jobs[ji];
        bk.R_max = std::max(bk.R_max, j.R);
        bk.n_tab_max = std::max(bk.n_tab_max, j.n_tab);
        bk.n_dense_max = std::max(bk.n_dense_max, j.n_dense);
    }
    const size_t col = (size_t)P.NH * P.T * P.Kp * 8;
    size_t o = 0;
    bk.o_alpha = o, o += al(col);
    bk.o_beta = o, o += al(col);
    bk.o_eG = o, o += al(col);
    bk.o_c = o, o += al((size_t)P.NH * P.T * 8);
    bk.o_W = o, o += al((size_t)P.T * P.Kp * 4);
    bk.o_Wc = o, o += al(P.rare_common ? (size_t)B->panel.Tc * P.Kp * 4 : 0);
    bk.o_tabs = o, o += al((size_t)std::max(bk.n_tab_max, 1) * sizeof(TabEnt));
    bk.o_dense = o, o += al((size_t)std::max(bk.n_dense_max, 1) * P.Kp * sizeof(TabEnt));
    bk.o_xprob = o, o += al((size_t)bk.R_max * 32);
    bk.o_snp_type = o, o += al(P.nSNPs);
    bk.o_rate = o, o += al((size_t)P.T * 8);
    bk.o_hapLocal = o, o += al(P.rare_common ? (size_t)P.nSNPs * 24 : 0);
    bk.o_blk = o, o += al(P.NH == 3 ? BlockScratch::bytes(P.T) : 0);
    // haplotype classes of the sweep kernel
    {
        const size_t KA = (size_t)bk.geo.NT * bk.geo.EPT, Tn = bk.classes ? (size_t)P.T : 0;
        bk.o_cinfo = o, o += al((Tn + 1) * 4);
        bk.o_cperm = o, o += al(Tn * KA * 2);
        bk.o_ccls = o, o += al(Tn * KA);
        bk.o_cent = o, o += al(Tn * bk.geo.NT);
        bk.o_crec = o, o += al(Tn * CLS_LANES * 32 * 16);
    }
    bk.slot_bytes = o;
    int cap = (g_sms / bk.geo.CL) * occ;
    const size_t by_mem = std::max<size_t>(1, *mem_budget / std::max<size_t>(bk.slot_bytes, 1));
    bk.n_slots = (int)std::min<size_t>(std::min<size_t>(cap, bk.jobs.size()), by_mem);
    CK(cached_alloc(bk.djobs, g_dcache_jobs, bk.jobs.size() * sizeof(JobDev)));
    CK(cached_alloc(bk.hjobs, g_hcache_jobs, bk.jobs.size() * sizeof(JobDev)));
    bk.perform_block = (P.flags & QUILT_F_PERFORM_BLOCK_GIBBS) != 0;
    bk.do_shard = (P.flags & QUILT_F_DO_SHARD_BLOCK_GIBBS) != 0;
    bk.debug = (P.flags & (QUILT_F_RETURN_ALPHA | QUILT_F_RETURN_EXTRA)) != 0;
    return QUILT_OK;
}

void make_jobdev(const QuiltGpuBatch* B, const Bucket& bk, const HostJob& j, int slot, JobDev* D) {
    std::memset(D, 0, sizeof(*D));
    char* s = (char*)bk.slots + (size_t)slot * bk.slot_bytes;
    const char* in = (const char*)B->din().p + j.in_off;
    char* out = (char*)B->dout().p + j.out_off;
    D->R = j.R;
    D->first_read = (j.a.flags & QUILT_F_GIBBS_INITIALIZE_AT_FIRST_READ) ? 0 : j.a.first_read_for_gibbs_initialization;
    D->n_dense = j.n_dense;
    D->alpha = (double*)(s + bk.o_alpha);
    D->beta = (double*)(s + bk.o_beta);
    D->eG = (double*)(s + bk.o_eG);
    D->c = (double*)(s + bk.o_c);
    D->W = (uint32_t*)(s + bk.o_W);
    D->Wc = (uint32_t*)(s + bk.o_Wc);
    D->desc = (ReadDesc*)(in + j.li.desc);
    D->tabs = (TabEnt*)(s + bk.o_tabs);
    D->dense = (TabEnt*)(s + bk.o_dense);
    D->xprob = (double*)(s + bk.o_xprob);
    D->snp_type = (uint8_t*)(s + bk.o_snp_type);
    D->rate = (double*)(s + bk.o_rate);
    D->hapLocal = (double*)(s + bk.o_hapLocal);
    D->cinfo = bk.classes ? (int32_t*)(s + bk.o_cinfo) : nullptr;
    D->cperm = (uint16_t*)(s + bk.o_cperm);
    D->ccls = (uint8_t*)(s + bk.o_ccls);
    D->cent = (uint8_t*)(s + bk.o_cent);
    D->crec = (uint4*)(s + bk.o_crec);
    D->which = (const int32_t*)(in + j.li.which);
    D->rs = (const int32_t*)(in + j.li.rs);
    D->roff = (const int32_t*)(in + j.li.roff);
    D->u = (const int32_t*)(in + j.li.u);
    D->pRA = (const double*)(in + j.li.pRA);
    D->wif0 = (const int32_t*)(in + j.li.wif0);
    D->ts = (const int32_t*)(in + j.li.ts);
    D->ginfo = (const int32_t*)(in + j.li.ginfo);
    D->dense_reads = (const int32_t*)(in + j.li.dense_reads);
    D->runif_reads = (const double*)(in + j.li.runif_reads);
    D->runif_shard = (const double*)(in + j.li.runif_shard);
    D->tm = (const double*)(in + j.li.tm);
    D->H0 = (const int32_t*)(in + j.li.H0);
    D->runif_block = (const double*)(in + j.li.runif_block);
    D->runif_H_class = (const double*)(in + j.li.runif_H_class);
    D->ep_stream = (j.a.unif_stream && !(j.a.flags & QUILT_F_SAMPLE_IS_DIPLOID)) ? 1 : 0;
    D->L_grid = (const int32_t*)(in + j.li.L_grid);
    D->blk = (unsigned char*)(s + bk.o_blk);
    D->H = (int32_t*)(out + j.lo.H);
    D->Hclass = (int32_t*)(out + j.lo.Hclass);
    D->lik = (double*)(out + j.lo.lik);
    D->underflow = (int32_t*)(out + j.lo.underflow);
    char* outB = (char*)B->doutB.p + j.outB_off;
    D->hapProbs = (double*)((j.lo.hap_dev_only ? outB : out) + j.lo.hap);
    D->genM = (double*)((j.lo.gen_dev_only ? outB : out) + j.lo.genM);
    D->genF = (double*)((j.lo.gen_dev_only ? outB : out) + j.lo.genF);
    D->cat_out = (int32_t*)(out + j.lo.cat);
    D->Hs = (int32_t*)(out + j.lo.Hs);
}

// grid = (ceil(R_max / 256), jobs): labels start from H0; read_category export
__global__ void __launch_bounds__(256) k_copy_H(const JobDev* __restrict__ jobs) {
    const JobDev& J = jobs[blockIdx.y];
    const int r = blockIdx.x * 256 + threadIdx.x;
    if (r < J.R) J.H[r] = J.H0[r];
    if (r == 0) J.lik[LIK_EP_POS] = 0.0;  // episode-stream position of the three-haplotype kernels
}
// labels after sampling sweep i (list_of_ending_read_labels.push_back(clone(H)), gibbs-nipt.cpp:3104); only launched when
// there is more than one sampling sweep
__global__ void __launch_bounds__(256) k_snapshot_H(const JobDev* __restrict__ jobs, int i) {
    const JobDev& J = jobs[blockIdx.y];
    const int r = blockIdx.x * 256 + threadIdx.x;
    if (r < J.R) J.Hs[(size_t)i * J.R + r] = J.H[r];
}
__global__ void __launch_bounds__(256) k_export_cat(const JobDev* __restrict__ jobs) {
    const JobDev& J = jobs[blockIdx.y];
    const int r = blockIdx.x * 256 + threadIdx.x;
    if (r < J.R) J.cat_out[r] = J.desc[r].cat;
}

int run_prep(QuiltGpuBatch* B, Bucket& bk, int n, const JobDev* dj) {
    const BatchParams& P = bk.P;
    const int kb = (P.Kp + 255) / 256;
    if (P.rare_common) {
        k_unpack_common<<<dim3(kb, B->panel.Tc, n), 256, 0, g_stream>>>(B->panel, dj, P.K, P.Kp, 1);
        LAUNCHED_K("k_unpack_common");
        k_assemble_all<<<dim3(kb, (P.T + ASM_GPB - 1) / ASM_GPB, n), 256, 0, g_stream>>>(B->panel, dj, P.K, P.Kp, P.T);
        LAUNCHED_K("k_assemble_all");
        k_scatter_rare<<<dim3((P.K + 255) / 256, n), 256, 0, g_stream>>>(B->panel, dj, P.K, P.Kp);
        LAUNCHED_K("k_scatter_rare");
        k_snp_type<<<dim3(P.T, n), 128, 0, g_stream>>>(B->panel, dj, P.K, P.Kp);
        LAUNCHED_K("k_snp_type");
    } else {
        k_unpack_common<<<dim3(kb, P.T, n), 256, 0, g_stream>>>(B->panel, dj, P.K, P.Kp, 0);
        LAUNCHED_K("k_unpack_common");
    }
    if (bk.classes) {
        const int KA = bk.geo.NT * bk.geo.EPT;
        const size_t csm = class_dyn_smem(P.Kp, KA);
        CK(cudaFuncSetAttribute(k_build_classes, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csm));
        k_build_classes<<<dim3(P.T, n), CLS_NT, csm, g_stream>>>(P, dj, bk.geo.NT, bk.geo.EPT);
        LAUNCHED_K("k_build_classes");
    }
    {
        const int tsm = 3 * P.Kp * 4;
        CK(cudaFuncSetAttribute(k_build_tables, cudaFuncAttributeMaxDynamicSharedMemorySize, tsm));
        k_build_tables<<<dim3(P.T, n), TAB_WARPS * 32, tsm, g_stream>>>(P, dj);
        LAUNCHED_K("k_build_tables");
    }
    if (bk.n_dense_max > 0) {
        k_build_dense<<<dim3(bk.n_dense_max, n), 256, 0, g_stream>>>(P, dj);
        LAUNCHED_K("k_build_dense");
    }
    CK(cudaGetLastError());
    return QUILT_OK;
}

template <int NT, int EPT>
int run_wave_t(QuiltGpuBatch* B, Bucket& bk, int n, const JobDev* dj, bool timed) {
    const BatchParams& P = bk.P;
    const int CL = bk.geo.CL;
    const SweepSmemLayout L = sweep_smem_layout(NT * EPT, P.NH, NT, CL);
    // (buckets of different K share a template instance: re-arm the opt-in shared-memory size for this one)
    if (CL == 2) {
        if constexpr (NT == 256 && EPT == 16) {
            CK(cudaFuncSetAttribute(k_sweep<NT, EPT, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
        } else {
            return set_err(QUILT_ERR_UNSUPPORTED, "no cluster sweep kernel for this geometry");
        }
    } else if (P.NH == 2 && bk.classes) {
        CK(cudaFuncSetAttribute(k_sweep<NT, EPT, 2, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    } else if (P.NH == 2) {
        CK(cudaFuncSetAttribute(k_sweep<NT, EPT, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    } else if constexpr (nipt_geo<NT, EPT>()) {
        CK(cudaFuncSetAttribute(k_sweep<NT, EPT, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    } else {
        return set_err(QUILT_ERR_UNSUPPORTED, "no three-haplotype sweep kernel for this K");
    }
    k_copy_H<<<dim3((bk.R_max + 255) / 256, n), 256, 0, g_stream>>>(dj);
    LAUNCHED_K("k_copy_H");
    int rc = run_prep(B, bk, n, dj);
    if (rc != QUILT_OK) return rc;
    if (P.flags & QUILT_F_GIBBS_INITIALIZE_ITERATIVELY) {
        k_init_iterative<<<dim3(P.T, n), 256, 0, g_stream>>>(P, dj);
        LAUNCHED_K("k_init_iterative");
    } else {
        k_make_eG<<<dim3(P.T, n), 256, 0, g_stream>>>(P, dj);
        LAUNCHED_K("k_make_eG");
        if (CL == 2)
            k_fb_generic<512, 16><<<dim3(n, P.NH), 512, 0, g_stream>>>(P, dj, 1);  // one CTA holds all K <= 8192 states here
        else
            k_fb_generic<NT, EPT><<<dim3(n, P.NH), NT, 0, g_stream>>>(P, dj, 1);
        LAUNCHED_K("k_fb_generic");
    }
    int episode = 0;
    for (int it = 0; it < P.n_its; it++) {
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (timed) {
            CK(cudaEventCreate(&e0));
            CK(cudaEventCreate(&e1));
            CK(cudaEventRecord(e0, g_stream));
        }
        bool blk_next = false;
        if (bk.perform_block)
            for (int b : bk.block_its) blk_next = blk_next || (b == it);
        // alpha columns are written only when a consumer follows this sweep (see k_sweep)
        const int store_alpha = (it >= P.n_burn || (P.NH == 3 && blk_next) || bk.debug || it == P.n_its - 1) ? 1 : 0;
        if (CL == 2) {
            if constexpr (NT == 256 && EPT == 16) CK(launch_cl(k_sweep<NT, EPT, 2, 2>, 2 * n, NT, (size_t)L.total, 2, P, dj, it, store_alpha));
        } else if (P.NH == 2 && bk.classes) {
            k_sweep<NT, EPT, 2, 1, true><<<n, NT, L.total, g_stream>>>(P, dj, it, store_alpha);
        } else if (P.NH == 2) {
            k_sweep<NT, EPT, 2><<<n, NT, L.total, g_stream>>>(P, dj, it, store_alpha);
        } else if constexpr (nipt_geo<NT, EPT>()) {
            k_sweep<NT, EPT, 3><<<n, NT, L.total, g_stream>>>(P, dj, it, store_alpha);
        }
        LAUNCHED_K("k_sweep");
        if (timed) {
            CK(cudaEventRecord(e1, g_stream));
            B->sweep_events.emplace_back(e0, e1);
        }
        bool blk = false;
        if (bk.perform_block)
            for (int b : bk.block_its) blk = blk || (b == it);
        if (blk) {
            // Diploid block resampler (gibbs-nipt-block.cpp:1636-1967): with two haplotypes the third column of the
            // reference's local matrices is zero, every permutation score is NaN and the identity is always kept
            // (DESIGN.md "block Gibbs, diploid"); the trailing backward passes reproduce beta unchanged.  Only the
            // shard pass alters state.
            if (P.NH == 3) {
                // three haplotypes: the real block resampler (block_nipt.cuh), then labels redrawn from H_class and the
                // whole state rebuilt (gibbs-nipt-block.cpp:1900-1958; the generic backward is overwritten by the fast one)
                if constexpr (nipt_geo<NT, EPT>()) {
                    if (P.T > 2) {
                        k_block_rate<<<dim3(P.T, n), 256, 0, g_stream>>>(P, dj);
                        LAUNCHED_K("k_block_rate");
                        {
                            // rate / peak list / availability of the block definition live in shared memory when they fit
                            const size_t bds = (size_t)P.T * 13 + 16;
                            const int in_smem = bds <= (size_t)160 * 1024;
                            if (in_smem && bds > 48 * 1024) CK(cudaFuncSetAttribute(k_block_define, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bds));
                            k_block_define<<<n, BD_NT, in_smem ? bds : 0, g_stream>>>(P, dj, in_smem);
                        }
                        LAUNCHED_K("k_block_define");
                        k_block_nipt<NT, EPT><<<n, NT, 0, g_stream>>>(P, dj, episode);
                        LAUNCHED_K("k_block_nipt");
                    }
                    k_sample_H<<<n, 256, 0, g_stream>>>(P, dj, episode);
                    LAUNCHED_K("k_sample_H");
                    k_make_eG<<<dim3(P.T, n), 256, 0, g_stream>>>(P, dj);
                    LAUNCHED_K("k_make_eG");
                    k_fb_generic<NT, EPT><<<dim3(n, P.NH), NT, 0, g_stream>>>(P, dj, 0);
                    LAUNCHED_K("k_fb_generic");
                    k_bwd_fast<NT, EPT><<<dim3(n, P.NH), NT, 0, g_stream>>>(P, dj);
                    LAUNCHED_K("k_bwd_fast");
                }
            }
            if (bk.do_shard && P.T > 1) {
                const size_t ssm = (size_t)3 * 2 * NT * EPT * 8;  // ring of three column pairs
                if (CL == 2) {
                    if constexpr (NT == 256 && EPT == 16) {
                        CK(cudaFuncSetAttribute(k_shard<NT, EPT, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssm));
                        CK(launch_cl(k_shard<NT, EPT, 2>, 2 * n, NT, ssm, 2, P, dj, episode));
                    }
                } else {
                    CK(cudaFuncSetAttribute(k_shard<NT, EPT, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssm));
                    k_shard<NT, EPT, 1><<<n, NT, ssm, g_stream>>>(P, dj, episode);
                }
                LAUNCHED_K("k_shard");
            }
            episode++;
        }
        if (it >= P.n_burn) {
            const int n_sample = P.n_its - P.n_burn;
            if (n_sample > 1) {
                k_snapshot_H<<<dim3((bk.R_max + 255) / 256, n), 256, 0, g_stream>>>(dj, it - P.n_burn);
                LAUNCHED_K("k_snapshot_H");
            }
            if (P.NH == 2)
                k_happrobs<2><<<dim3(P.T, n), 256, 0, g_stream>>>(P, dj, it == P.n_burn, it == P.n_its - 1, 1.0 / double(n_sample));
            else
                k_happrobs<3><<<dim3(P.T, n), 384, 0, g_stream>>>(P, dj, it == P.n_burn, it == P.n_its - 1, 1.0 / double(n_sample));
            LAUNCHED_K("k_happrobs");
        }
    }
    k_export_cat<<<dim3((bk.R_max + 255) / 256, n), 256, 0, g_stream>>>(dj);
    LAUNCHED_K("k_export_cat");
    CK(cudaGetLastError());
    return QUILT_OK;
}

int fetch_debug(QuiltGpuBatch* B, Bucket& bk, int w0, int n) {
    const BatchParams& P = bk.P;
    CK(cudaStreamSynchronize(g_stream));
    const size_t cols = (size_t)P.NH * P.T * P.Kp;
    std::vector<double> tmp(cols);
    for (int i = 0; i < n; i++) {
        HostJob& j = B->jobs[bk.jobs[w0 + i]];
        const char* s = (const char*)bk.slots + (size_t)i * bk.slot_bytes;
        if (P.flags & QUILT_F_RETURN_ALPHA) {
            auto grab = [&](size_t off, std::vector<double>& dst) -> int {
                CK(cudaMemcpy(tmp.data(), s + off, cols * 8, cudaMemcpyDeviceToHost));
                dst.assign((size_t)P.NH * P.T * P.K, 0.0);
                for (int h = 0; h < P.NH; h++)
                    for (int g = 0; g < P.T; g++)
                        std::memcpy(&dst[((size_t)h * P.T + g) * P.K], &tmp[((size_t)h * P.T + g) * P.Kp], (size_t)P.K * 8);
                return QUILT_OK;
            };
            int rc;
            if ((rc = grab(bk.o_alpha, j.dbg_alpha)) != QUILT_OK) return rc;
            if ((rc = grab(bk.o_beta, j.dbg_beta)) != QUILT_OK) return rc;
            if ((rc = grab(bk.o_eG, j.dbg_eG)) != QUILT_OK) return rc;
            j.dbg_c.assign((size_t)P.NH * P.T, 0.0);
            CK(cudaMemcpy(j.dbg_c.data(), s + bk.o_c, (size_t)P.NH * P.T * 8, cudaMemcpyDeviceToHost));
        }
        if (P.flags & QUILT_F_RETURN_EXTRA) {
            DBuf d;
            CK(d.alloc((size_t)P.K * j.R * 8));
            const JobDev* dj = (const JobDev*)bk.djobs.p + w0 + i;
            k_expand_eMatRead<<<j.R, 256, 0, g_stream>>>(P, dj, (double*)d.p);
            LAUNCHED_K("k_expand_eMatRead");
            CK(cudaStreamSynchronize(g_stream));
            j.dbg_eMatRead.assign((size_t)P.K * j.R, 0.0);
            CK(cudaMemcpy(j.dbg_eMatRead.data(), d.p, (size_t)P.K * j.R * 8, cudaMemcpyDeviceToHost));
        }
    }
    return QUILT_OK;
}

// JobDev records of every wave of a bucket (slot = position inside the wave), uploaded once
int upload_jobdevs(QuiltGpuBatch* B, Bucket& bk) {
    const int nj = (int)bk.jobs.size();
    JobDev* hj = (JobDev*)bk.hjobs.p;
    for (int q = 0; q < nj; q++) make_jobdev(B, bk, B->jobs[bk.jobs[q]], q % bk.n_slots, &hj[q]);
    CK(cudaMemcpyAsync(bk.djobs.p, hj, (size_t)nj * sizeof(JobDev), cudaMemcpyHostToDevice, g_stream));
    return QUILT_OK;
}

int run_wave(QuiltGpuBatch* B, Bucket& bk, int w0, int n, bool timed, bool prep_only) {
    if (bk.P.rare_common) {
        for (int i = 0; i < n; i++)
            CK(cudaMemsetAsync((char*)bk.slots + (size_t)i * bk.slot_bytes + bk.o_hapLocal, 0, (size_t)bk.P.nSNPs * 24, g_stream));
    }
    const JobDev* dj = (const JobDev*)bk.djobs.p + w0;
    int rc;
    if (prep_only) {
        k_copy_H<<<dim3((bk.R_max + 255) / 256, n), 256, 0, g_stream>>>(dj);
        LAUNCHED_K("k_copy_H");
        rc = run_prep(B, bk, n, dj);
        if (rc == QUILT_OK) {
            k_export_cat<<<dim3((bk.R_max + 255) / 256, n), 256, 0, g_stream>>>(dj);
            LAUNCHED_K("k_export_cat");
        }
    } else {
        rc = with_geo(bk.geo, [&](auto nt, auto ept) { return run_wave_t<decltype(nt)::value, decltype(ept)::value>(B, bk, n, dj, timed); });
    }
    if (rc != QUILT_OK) return rc;
    if (bk.debug) {
        rc = fetch_debug(B, bk, w0, n);
        if (rc != QUILT_OK) return rc;
    }
    return QUILT_OK;
}

// the shared wave-state arena is large enough for this batch and its buckets point into it
int ensure_slots(QuiltGpuBatch* B) {
    if (g_slots.bytes < B->slot_arena_bytes) {
        CK(cudaStreamSynchronize(g_stream));  // nothing may still be running in the old arena
        g_slots.release();
        CK(g_slots.alloc(B->slot_arena_bytes));
    }
    for (auto& bk : B->buckets) bk->slots = (char*)g_slots.p;
    return QUILT_OK;
}

int run_bucket(QuiltGpuBatch* B, Bucket& bk, bool timed, bool prep_only = false) {
    const int nj = (int)bk.jobs.size();
    int rc = upload_jobdevs(B, bk);
    if (rc != QUILT_OK) return rc;
    for (int w0 = 0; w0 < nj; w0 += bk.n_slots) {
        rc = run_wave(B, bk, w0, std::min(bk.n_slots, nj - w0), timed, prep_only);
        if (rc != QUILT_OK) return rc;
    }
    return QUILT_OK;
}

// per_it_likelihoods row from the device record (add_to_per_it_likelihoods, gibbs-nipt.cpp:1583-1621;
// calculate_likelihoods_values :1483-1519; rcpp_get_log_p_H_class gibbs-nipt-block.cpp:146-164)
void fill_lik_row(const QuiltGibbsArgs& a, const double* rec, int iteration, int n_rows, double* out) {
    const bool diploid = (a.flags & QUILT_F_SAMPLE_IS_DIPLOID) != 0;
    const double ff = a.ff;
    const double inf = std::numeric_limits<double>::infinity();
    const double d1 = rec[0], d2 = rec[1], d3 = diploid ? inf : rec[2];  // diploid: c3 is all zero -> -log(0)
    const double prior[3] = {0.5, (1 - ff) / 2, (ff / 2)};
    double dH = 0;
    for (int h = 0; h < 3; h++)
        if (rec[3 + h] > 0) dH += rec[3 + h] * std::log(prior[h]);
    const int burn = a.n_gibbs_burn_in_its;
    const int i_result_it = (iteration + 1 > burn) ? iteration - burn : -2;
    auto at = [&](int col) -> double& { return out[(size_t)col * n_rows + iteration]; };
    at(0) = 1;
    at(1) = 1;
    at(2) = iteration + 1;
    at(3) = i_result_it + 1;
    at(4) = d1;
    at(5) = d2;
    at(6) = d3;
    at(7) = d1 + d2 + d3;
    at(8) = dH;
    at(9) = at(7) + dH;
    const double rc[3] = {rec[3], rec[4], rec[5]};
    const int n = (int)(rc[0] + rc[1] + rc[2]);
    double r = lgamma_ts(1.0 * (n + 1.0));
    for (int i = 0; i < 3; i++)
        if (prior[i] > 0) r += rc[i] * std::log(prior[i]) - lgamma_ts(1.0 * (rc[i] + 1.0));
    at(10) = r;
    at(11) = 1;
    double lp = 0;
    if (a.flags & QUILT_F_RECORD_READ_SET) {
        const double vals[8] = {0, std::log(0.5), std::log(0.5 - ff * 0.5), std::log(ff * 0.5), std::log(1.0 - ff * 0.5), std::log(0.5 + ff * 0.5),
                                std::log(0.5), 0};
        for (int q = 0; q < 8; q++)
            if (rec[6 + q] > 0) lp += rec[6 + q] * vals[q];
    }
    at(12) = lp;
}

}  // namespace

// ================================================================================================== C ABI
extern "C" {

int quilt_gpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int quilt_gpu_set_device(int32_t device) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_stream && device != g_device) return set_err(QUILT_ERR_BAD_ARG, "device already initialised");
    g_device = device;
    return ensure_device();
}

const char* quilt_gpu_last_error(void) { return g_err.c_str(); }

int64_t quilt_gpu_kernel_launches(void) { return g_launches.load(); }

int quilt_gpu_section_timing(int32_t enable) {
    int rc = ensure_device();
    if (rc != QUILT_OK) return rc;
    CK(cudaStreamSynchronize(g_stream));
    std::lock_guard<std::mutex> lk(g_sect.mu);
    g_sect.reset();
    g_sect.on = enable != 0;
    if (g_sect.on) {
        CK(cudaEventCreate(&g_sect.begin));
        CK(cudaEventRecord(g_sect.begin, g_stream));
    }
    return QUILT_OK;
}

int64_t quilt_gpu_section_report(char* buf, int64_t cap) {
    if (!g_stream) return 0;
    cudaStreamSynchronize(g_stream);
    std::lock_guard<std::mutex> lk(g_sect.mu);
    std::vector<std::pair<std::string, std::pair<double, long>>> acc;  // first-appearance order
    cudaEvent_t prev = g_sect.begin;
    double total = 0;
    for (auto& m : g_sect.marks) {
        float ms = 0;
        if (prev && cudaEventElapsedTime(&ms, prev, m.second) == cudaSuccess) {
            auto it = std::find_if(acc.begin(), acc.end(), [&](const auto& a) { return a.first == m.first; });
            if (it == acc.end()) {
                acc.push_back({m.first, {0.0, 0}});
                it = acc.end() - 1;
            }
            it->second.first += ms;
            it->second.second += 1;
            total += ms;
        }
        prev = m.second;
    }
    std::string out = "section                      launches     total ms       avg ms    share\n";
    char line[160];
    for (auto& a : acc) {
        std::snprintf(line, sizeof line, "%-28s %8ld %12.3f %12.4f %7.1f%%\n", a.first.c_str(), a.second.second, a.second.first,
                      a.second.first / std::max<long>(a.second.second, 1), total > 0 ? 100.0 * a.second.first / total : 0.0);
        out += line;
    }
    std::snprintf(line, sizeof line, "%-28s %8zu %12.3f\n", "TOTAL", g_sect.marks.size(), total);
    out += line;
    if (buf && cap > 0) {
        const size_t n = std::min<size_t>(out.size(), (size_t)cap - 1);
        std::memcpy(buf, out.data(), n);
        buf[n] = 0;
    }
    return (int64_t)out.size() + 1;
}

void quilt_gpu_release_panel_cache(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_panels.clear();
    g_dcache_in.clear();
    g_dcache_out.clear();
    g_slots.release();
    g_hcache_in.clear();
    g_hcache_out.clear();
    g_dcache_outB.clear();
    g_dcache_jobs.clear();
    g_dcache_hap.clear();
    g_hcache_jobs.clear();
}

int quilt_gpu_batch_free(QuiltGpuBatch* b) {
    if (b) {
        if (g_stream) cudaStreamSynchronize(g_stream);
        delete b;
    }
    return QUILT_OK;
}

static int stage_impl(int32_t n, const QuiltGibbsArgs* args, QuiltGpuBatch** batch, bool upload) {
    if (n <= 0 || !args || !batch) return set_err(QUILT_ERR_BAD_ARG, "bad batch arguments");
    int rc = ensure_device();
    if (rc != QUILT_OK) return rc;
    std::unique_ptr<QuiltGpuBatch> B(new QuiltGpuBatch());
    B->n = n;
    B->jobs.resize(n);
    for (int i = 0; i < n; i++) {
        if ((rc = validate(args[i])) != QUILT_OK) return rc;
        if (args[i].panel != args[0].panel && !same_panel_struct(*args[i].panel, *args[0].panel))
            return set_err(QUILT_ERR_UNSUPPORTED, "all calls of a batch must share one panel");
        B->jobs[i].a = args[i];
    }
    if ((rc = get_panel(args[0].panel, &B->panel, &B->panel_ref)) != QUILT_OK) return rc;
    std::map<BucketKey, int> index;
    size_t in_total = 0, out_total = 0, outB_total = 0;
    {
        // descriptor building is O(sum J) per job and independent across jobs
        std::vector<int> rcs(n, QUILT_OK);
        std::vector<std::string> errs(n);
        QuiltGpuBatch* Bp = B.get();
        parallel_for(n, [&](int i) {
            rcs[i] = prepare_job(Bp->jobs[i]);
            if (rcs[i] != QUILT_OK) errs[i] = g_err;
        });
        for (int i = 0; i < n; i++)
            if (rcs[i] != QUILT_OK) return set_err(rcs[i], errs[i]);
    }
    for (int i = 0; i < n; i++) {
        HostJob& j = B->jobs[i];
        layout_in(j);
        const BucketKey key = bucket_key(j.a);
        auto it = index.find(key);
        if (it == index.end()) {
            auto bk = std::make_unique<Bucket>();
            make_params(j.a, &bk->P);
            bk->block_its.assign(j.a.block_gibbs_iterations, j.a.block_gibbs_iterations + j.a.n_block_gibbs_iterations);
            index[key] = (int)B->buckets.size();
            j.bucket = (int)B->buckets.size();
            B->buckets.push_back(std::move(bk));
        } else {
            j.bucket = it->second;
        }
        B->buckets[j.bucket]->jobs.push_back(i);
    }
    // arena offsets in (bucket, position) order: the jobs of a wave are contiguous, so a wave's inputs / outputs move
    // with one copy each (pipelined path)
    for (auto& bk : B->buckets)
        for (int ji : bk->jobs) {
            HostJob& j = B->jobs[ji];
            j.in_off = in_total;
            j.out_off = out_total;
            j.outB_off = outB_total;
            in_total += j.li.end;
            out_total += j.lo.end;
            outB_total += j.lo.endB;
        }
    B->in_bytes = in_total;
    B->out_bytes = out_total;
    B->outB_bytes = outB_total;
    CK(cached_alloc(B->doutB, g_dcache_outB, outB_total));
    {
        cudaError_t e;
        B->din_ = g_dcache_in.acquire(in_total, &e);
        CK(e);
        B->dout_ = g_dcache_out.acquire(out_total, &e);
        CK(e);
        B->hin_ = g_hcache_in.acquire(in_total, &e);
        CK(e);
        B->hout_ = g_hcache_out.acquire(out_total, &e);
        CK(e);
    }
    if (upload) {
        QuiltGpuBatch* Bp = B.get();
        char* hin = (char*)Bp->hin().p;
        parallel_for(n, [&](int i) { fill_in(Bp->jobs[i], hin + Bp->jobs[i].in_off); });
        CK(cudaMemcpyAsync(B->din().p, B->hin().p, in_total, cudaMemcpyHostToDevice, g_stream));
    }
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    free_b += g_slots.bytes;  // the shared wave-state arena is reused (or regrown) by this batch
    size_t budget = (size_t)(free_b * 0.94);
    size_t arena = 0;
    for (auto& bk : B->buckets) {
        size_t mb = budget;
        if ((rc = setup_bucket(B.get(), *bk, &mb)) != QUILT_OK) return rc;
        arena = std::max(arena, (size_t)bk->n_slots * bk->slot_bytes);
    }
    B->slot_arena_bytes = arena;
    CK(cudaEventCreate(&B->ev0));
    CK(cudaEventCreate(&B->ev1));
    if (upload) CK(cudaStreamSynchronize(g_stream));  // (the pipelined paths prepare a batch while earlier waves still run)
    *batch = B.release();
    return QUILT_OK;
}

int quilt_gpu_batch_stage(int32_t n, const QuiltGibbsArgs* args, QuiltGpuBatch** batch) {
    std::lock_guard<std::mutex> lk(g_mu);
    return stage_impl(n, args, batch, true);
}

int quilt_gpu_batch_run(QuiltGpuBatch* B) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!B) return set_err(QUILT_ERR_BAD_ARG, "null batch");
    for (auto& p : B->sweep_events) {
        cudaEventDestroy(p.first);
        cudaEventDestroy(p.second);
    }
    B->sweep_events.clear();
    {
        int rc = ensure_slots(B);
        if (rc != QUILT_OK) return rc;
    }
    CK(cudaMemsetAsync(B->dout().p, 0, B->out_bytes, g_stream));
    if (B->outB_bytes) CK(cudaMemsetAsync(B->doutB.p, 0, B->outB_bytes, g_stream));
    CK(cudaEventRecord(B->ev0, g_stream));
    for (auto& bk : B->buckets) {
        int rc = run_bucket(B, *bk, true);
        if (rc != QUILT_OK) return rc;
    }
    CK(cudaEventRecord(B->ev1, g_stream));
    B->ran = true;
    B->fetched_raw = false;
    return QUILT_OK;
}

int quilt_gpu_batch_sync(QuiltGpuBatch* B) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!B) return set_err(QUILT_ERR_BAD_ARG, "null batch");
    CK(cudaStreamSynchronize(g_stream));
    if (B->ran) {
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, B->ev0, B->ev1));
        B->total_ms = ms;
        double s = 0;
        for (auto& p : B->sweep_events) {
            CK(cudaEventElapsedTime(&ms, p.first, p.second));
            s += ms;
        }
        B->sweep_ms = s;
        B->n_sweep_launches = (int)B->sweep_events.size();
    }
    return QUILT_OK;
}

int quilt_gpu_batch_timing(QuiltGpuBatch* B, double* total_ms, double* sweep_ms, int32_t* n_sweep_launches) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!B) return set_err(QUILT_ERR_BAD_ARG, "null batch");
    if (total_ms) *total_ms = B->total_ms;
    if (sweep_ms) *sweep_ms = B->sweep_ms;
    if (n_sweep_launches) *n_sweep_launches = B->n_sweep_launches;
    return QUILT_OK;
}

int quilt_gpu_batch_bytes(QuiltGpuBatch* B, int64_t* h2d, int64_t* d2h, double* sweep_alg) {
    if (!B) return set_err(QUILT_ERR_BAD_ARG, "null batch");
    if (h2d) *h2d = (int64_t)B->in_bytes;
    if (d2h) *d2h = (int64_t)B->out_bytes;
    if (sweep_alg) {
        double s = 0;
        for (const HostJob& j : B->jobs) {
            const int NH = (j.a.flags & QUILT_F_SAMPLE_IS_DIPLOID) ? 2 : 3;
            s += (double)j.n_its * 8.0 * j.a.K * (5.0 * NH * j.a.nGrids + j.R);
        }
        *sweep_alg = s;
    }
    return QUILT_OK;
}

static void unpack_job(const QuiltGpuBatch* B, int i, QuiltGibbsOut* out) {
    const HostJob& j = B->jobs[i];
    const QuiltGibbsArgs& a = j.a;
    QuiltGibbsOut& o = out[i];
    const char* base = (const char*)B->hout().p + j.out_off;
    const int under = *reinterpret_cast<const int32_t*>(base + j.lo.underflow);
    o.underflow_problem = under ? 1 : 0;
    {
        // first sweep whose underflow check failed (gibbs-nipt.cpp:2959-2969) and, in episode-stream mode, how many
        // stream values the reference would have drawn: the episodes of earlier sweeps only
        const double* lik = reinterpret_cast<const double*>(base + j.lo.lik);
        int it_u = -1;
        if (under)
            for (int it = 0; it < j.n_its && it_u < 0; it++)
                if (lik[(size_t)it * LIK_N + 14] != 0.0) it_u = it;
        o.underflow_iteration = it_u;
        o.n_unif_consumed = 0;
        if (a.unif_stream && (a.flags & QUILT_F_PERFORM_BLOCK_GIBBS)) {
            if (a.flags & QUILT_F_SAMPLE_IS_DIPLOID) {
                int n_done = 0;
                for (int e = 0; e < a.n_block_gibbs_iterations; e++) {
                    const int b = a.block_gibbs_iterations[e];
                    if (b >= 0 && b < j.n_its && (it_u < 0 || b < it_u)) n_done++;
                }
                o.n_unif_consumed = (int64_t)diploid_episode_base(a, n_done);
            } else {
                o.n_unif_consumed = (int64_t)lik[LIK_EP_POS];
            }
        }
    }
    const size_t n3 = (size_t)a.nSNPs * 3;
    if (!j.lo.hap_dev_only) {
        const double* hp = reinterpret_cast<const double*>(base + j.lo.hap);
        if (o.hapProbs_t) std::memcpy(o.hapProbs_t, hp, n3 * 8);
        if (!j.lo.gen_dev_only) {
            if (o.genProbsM_t) std::memcpy(o.genProbsM_t, base + j.lo.genM, n3 * 8);
            if (o.genProbsF_t) std::memcpy(o.genProbsF_t, base + j.lo.genF, n3 * 8);
        } else if (!under) {
            // one sampling sweep: genProbs are the products of the haplotype probabilities, in the reference's own
            // expression order (gibbs-small.cpp:621-633; device twin: k_happrobs) — bit-identical to what the device holds
            // (the common-SNP variant also fills genProbsF_t for diploid samples, from a zero third haplotype, :627-633; the
            //  rare/common variant leaves it zero there, :838-846)
            const bool nipt = !(a.flags & QUILT_F_SAMPLE_IS_DIPLOID) || !(a.flags & QUILT_F_MAKE_EMATREAD_RARE_COMMON);
            for (int s = 0; s < a.nSNPs; s++) {
                const double h0 = hp[3 * (size_t)s], h1 = hp[3 * (size_t)s + 1], h2 = hp[3 * (size_t)s + 2];
                if (o.genProbsM_t) {
                    double* g = o.genProbsM_t + 3 * (size_t)s;
                    g[0] = (1 - h0) * (1 - h1);
                    g[1] = h0 * (1 - h1) + h1 * (1 - h0);
                    g[2] = h0 * h1;
                }
                if (o.genProbsF_t) {
                    double* g = o.genProbsF_t + 3 * (size_t)s;
                    if (nipt) {
                        g[0] = (1 - h0) * (1 - h2);
                        g[1] = h0 * (1 - h2) + h2 * (1 - h0);
                        g[2] = h0 * h2;
                    } else {
                        g[0] = g[1] = g[2] = 0.0;
                    }
                }
            }
        } else {
            if (o.genProbsM_t) std::memset(o.genProbsM_t, 0, n3 * 8);
            if (o.genProbsF_t) std::memset(o.genProbsF_t, 0, n3 * 8);
        }
    }
    if (o.H) std::memcpy(o.H, base + j.lo.H, (size_t)j.R * 4);
    if (o.H_sample_its && a.n_gibbs_sample_its > 0)
        std::memcpy(o.H_sample_its, base + (a.n_gibbs_sample_its > 1 ? j.lo.Hs : j.lo.H), (size_t)a.n_gibbs_sample_its * j.R * 4);
    if (o.H_class && (a.flags & QUILT_F_RECORD_READ_SET)) std::memcpy(o.H_class, base + j.lo.Hclass, (size_t)j.R * 4);
    if (o.read_category) std::memcpy(o.read_category, base + j.lo.cat, (size_t)j.R * 4);
    if (o.per_it_likelihoods) {
        const int n_rows = (a.n_gibbs_sample_its == 0) ? 1 : j.n_its;
        std::memset(o.per_it_likelihoods, 0, (size_t)n_rows * 13 * 8);
        const double* lik = reinterpret_cast<const double*>(base + j.lo.lik);
        for (int it = 0; it < std::min(n_rows, j.n_its); it++) {
            fill_lik_row(a, lik + (size_t)it * LIK_N, it, n_rows, o.per_it_likelihoods);
        }
    }
    const int NH = (a.flags & QUILT_F_SAMPLE_IS_DIPLOID) ? 2 : 3;
    if (a.flags & QUILT_F_RETURN_ALPHA) {
        const size_t per = (size_t)a.nGrids * a.K;
        for (int h = 0; h < NH; h++) {
            if (o.alphaHat_t[h] && !j.dbg_alpha.empty()) std::memcpy(o.alphaHat_t[h], &j.dbg_alpha[h * per], per * 8);
            if (o.betaHat_t[h] && !j.dbg_beta.empty()) std::memcpy(o.betaHat_t[h], &j.dbg_beta[h * per], per * 8);
            if (o.eMatGrid_t[h] && !j.dbg_eG.empty()) std::memcpy(o.eMatGrid_t[h], &j.dbg_eG[h * per], per * 8);
            if (o.c[h] && !j.dbg_c.empty()) std::memcpy(o.c[h], &j.dbg_c[(size_t)h * a.nGrids], (size_t)a.nGrids * 8);
        }
    }
    if ((a.flags & QUILT_F_RETURN_EXTRA) && o.eMatRead_t && !j.dbg_eMatRead.empty())
        std::memcpy(o.eMatRead_t, j.dbg_eMatRead.data(), j.dbg_eMatRead.size() * 8);
}

int quilt_gpu_batch_fetch(QuiltGpuBatch* B, QuiltGibbsOut* out) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!B || !out) return set_err(QUILT_ERR_BAD_ARG, "null batch / out");
    if (!B->ran) return set_err(QUILT_ERR_BAD_ARG, "batch has not been run");
    if (!B->fetched_raw) {
        CK(cudaMemcpyAsync(B->hout().p, B->dout().p, B->out_bytes, cudaMemcpyDeviceToHost, g_stream));
        CK(cudaStreamSynchronize(g_stream));
        B->fetched_raw = true;
    }
    parallel_for(B->n, [&](int i) { unpack_job(B, i, out); });
    return QUILT_OK;
}

namespace {
// device-resident haplotype re-selection for a subset of jobs (defined with the selection code below): job ids[i] of `next`
// receives the list selected from job ids[i] of `prev`; asynchronous on the library stream
struct ChainSpec {
    QuiltGpuBatch* prev = nullptr;
    int nIndices = 4, L = 3, M = 1;
    const double* pad_unif = nullptr;  // [n_jobs x Ksubset], row = job id
};
int chain_select_subset(const ChainSpec& cs, QuiltGpuBatch* next, const int* ids, int m);
int chain_check(const ChainSpec& cs, QuiltGpuBatch* next);
}  // namespace

// Host buffers in, host buffers out.  Waves are pipelined: while the kernels of wave w run, the host fills the pinned
// input block of wave w + 1 (one H2D copy per wave on a copy stream) and unpacks the results of wave w - 1 (one D2H copy
// per wave on a second copy stream), so staging overlaps compute instead of preceding / following it.  With a chain
// specification every wave's haplotype lists are selected on the device from the previous batch's hapProbs_t right after
// the wave's inputs have landed; with `kept` the batch (device results included) survives the call for the next link.
namespace {
// One pipeline over the waves of one OR SEVERAL batches (the stages of a call chain): per wave host fill -> H2D (copy
// stream) -> [device-side selection of the haplotype lists] -> kernels -> D2H (copy stream) -> host unpack of the
// PREVIOUS wave.  Because the waves of consecutive stages are one sequence, the job preparation of stage s + 1 and the
// unpacking of stage s's last wave run while the GPU is busy with stage s / s + 1.
struct WavePipe {
    struct Wave {
        QuiltGpuBatch* B;
        QuiltGibbsOut* out;
        Bucket* bk;
        int w0, n;
        cudaEvent_t ev_in = nullptr, ev_done = nullptr, ev_out = nullptr;
    };
    std::vector<Wave> waves;
    size_t unpacked = 0;
    double t_fill = 0, t_unpack = 0;  // host time spent staging inputs / waiting for and unpacking results (trace)
    ~WavePipe() {
        for (auto& w : waves) {
            if (w.ev_in) cudaEventDestroy(w.ev_in);
            if (w.ev_done) cudaEventDestroy(w.ev_done);
            if (w.ev_out) cudaEventDestroy(w.ev_out);
        }
    }
    int unpack_through(size_t upto) {  // waves [unpacked, upto)
        const double t0 = now_ms();
        for (; unpacked < upto; unpacked++) {
            Wave& w = waves[unpacked];
            CK(cudaEventSynchronize(w.ev_out));
            parallel_for(w.n, [&](int i) { unpack_job(w.B, w.bk->jobs[w.w0 + i], w.out); });
        }
        t_unpack += now_ms() - t0;
        return QUILT_OK;
    }
    int push_batch(QuiltGpuBatch* B, QuiltGibbsOut* out, const ChainSpec* chain) {
        int rc = ensure_slots(B);
        if (rc != QUILT_OK) return rc;
        CK(cudaMemsetAsync(B->dout().p, 0, B->out_bytes, g_stream));
        if (B->outB_bytes) CK(cudaMemsetAsync(B->doutB.p, 0, B->outB_bytes, g_stream));
        for (auto& bk : B->buckets) {
            bool uploaded = false;
            for (int w0 = 0; w0 < (int)bk->jobs.size(); w0 += bk->n_slots) {
                waves.push_back(Wave{B, out, bk.get(), w0, std::min(bk->n_slots, (int)bk->jobs.size() - w0)});
                Wave& w = waves.back();
                const HostJob& first = B->jobs[w.bk->jobs[w.w0]];
                const HostJob& last = B->jobs[w.bk->jobs[w.w0 + w.n - 1]];
                const size_t in_lo = first.in_off, in_hi = last.in_off + last.li.end;
                const size_t out_lo = first.out_off, out_hi = last.out_off + last.lo.end;
                // host: this wave's inputs into the pinned block (runs while the previous wave computes)
                char* hin = (char*)B->hin().p;
                const double t_f = now_ms();
                parallel_for(w.n, [&](int i) {
                    const HostJob& j = B->jobs[w.bk->jobs[w.w0 + i]];
                    fill_in(j, hin + j.in_off);
                });
                t_fill += now_ms() - t_f;
                CK(cudaEventCreateWithFlags(&w.ev_in, cudaEventDisableTiming));
                CK(cudaEventCreateWithFlags(&w.ev_done, cudaEventDisableTiming));
                CK(cudaEventCreateWithFlags(&w.ev_out, cudaEventDisableTiming));
                CK(cudaMemcpyAsync((char*)B->din().p + in_lo, hin + in_lo, in_hi - in_lo, cudaMemcpyHostToDevice, g_copy_in));
                CK(cudaEventRecord(w.ev_in, g_copy_in));
                CK(cudaStreamWaitEvent(g_stream, w.ev_in, 0));
                if (!uploaded) {
                    if ((rc = upload_jobdevs(B, *w.bk)) != QUILT_OK) return rc;
                    uploaded = true;
                }
                if (chain && (rc = chain_select_subset(*chain, B, w.bk->jobs.data() + w.w0, w.n)) != QUILT_OK) return rc;
                if ((rc = run_wave(B, *w.bk, w.w0, w.n, false, false)) != QUILT_OK) return rc;
                CK(cudaEventRecord(w.ev_done, g_stream));
                CK(cudaStreamWaitEvent(g_copy_out, w.ev_done, 0));
                CK(cudaMemcpyAsync((char*)B->hout().p + out_lo, (const char*)B->dout().p + out_lo, out_hi - out_lo, cudaMemcpyDeviceToHost, g_copy_out));
                CK(cudaEventRecord(w.ev_out, g_copy_out));
                // host: results of the waves before this one (their D2H finishes while this wave computes)
                if ((rc = unpack_through(waves.size() - 1)) != QUILT_OK) return rc;
            }
        }
        B->ran = true;
        B->chained = chain != nullptr;
        return QUILT_OK;
    }
    int drain() {
        int rc = unpack_through(waves.size());
        if (rc != QUILT_OK) return rc;
        CK(cudaGetLastError());
        return QUILT_OK;
    }
};
void sync_all_streams() {
    cudaStreamSynchronize(g_stream);
    cudaStreamSynchronize(g_copy_in);
    cudaStreamSynchronize(g_copy_out);
}
}  // namespace

static int gibbs_batch_impl(int32_t n, const QuiltGibbsArgs* args, QuiltGibbsOut* out, const ChainSpec* chain, QuiltGpuBatch** kept) {
    if (!out) return set_err(QUILT_ERR_BAD_ARG, "null out");
    QuiltGpuBatch* B = nullptr;
    int rc = stage_impl(n, args, &B, false);
    if (rc != QUILT_OK) return rc;
    if (chain) rc = chain_check(*chain, B);
    {
        WavePipe pipe;
        if (rc == QUILT_OK) rc = pipe.push_batch(B, out, chain);
        if (rc == QUILT_OK) rc = pipe.drain();
        sync_all_streams();
    }
    if (kept && rc == QUILT_OK)
        *kept = B;
    else
        delete B;
    return rc;
}

int quilt_gpu_gibbs_batch(int32_t n, const QuiltGibbsArgs* args, QuiltGibbsOut* out) {
    std::lock_guard<std::mutex> lk(g_mu);
    return gibbs_batch_impl(n, args, out, nullptr, nullptr);
}

int quilt_gpu_gibbs_batch_chained(int32_t n, const QuiltGibbsArgs* args, QuiltGibbsOut* out, QuiltGpuBatch* prev, int32_t nIndices, int32_t L, int32_t M,
                                  const double* pad_unif, QuiltGpuBatch** kept) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (kept) *kept = nullptr;
    if (!prev) return gibbs_batch_impl(n, args, out, nullptr, kept);
    if (!pad_unif) return set_err(QUILT_ERR_BAD_ARG, "pad_unif is NULL");
    ChainSpec cs;
    cs.prev = prev;
    cs.nIndices = nIndices;
    cs.L = L;
    cs.M = M;
    cs.pad_unif = pad_unif;
    return gibbs_batch_impl(n, args, out, &cs, kept);
}

// The whole call chain of a set of (sample, chain) pairs in one call: stage s holds call s of every pair; from stage 1 on the
// haplotype lists are selected on the device from stage s - 1 (pad_unif[s - 1]: [n x Ksubset]).  One wave pipeline spans all
// stages, so preparing stage s + 1 on the host and unpacking stage s overlap the kernels.
int quilt_gpu_gibbs_chain(int32_t n_stages, int32_t n, const QuiltGibbsArgs* const* args, QuiltGibbsOut* const* out, int32_t nIndices, int32_t L,
                          int32_t M, const double* const* pad_unif) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (n_stages < 1 || n < 1 || !args || !out) return set_err(QUILT_ERR_BAD_ARG, "bad chain arguments");
    std::vector<std::unique_ptr<QuiltGpuBatch>> Bs;
    int rc = QUILT_OK;
    {
        WavePipe pipe;
        for (int s = 0; s < n_stages && rc == QUILT_OK; s++) {
            if (!args[s] || !out[s] || (s > 0 && (!pad_unif || !pad_unif[s - 1]))) {
                rc = set_err(QUILT_ERR_BAD_ARG, "null stage arguments");
                break;
            }
            QuiltGpuBatch* B = nullptr;
            const double t_a = now_ms();
            rc = stage_impl(n, args[s], &B, false);  // host preparation: the GPU is busy with the previous stage meanwhile
            if (rc != QUILT_OK) break;
            if (trace_on()) std::fprintf(stderr, "[quilt trace] chain stage %d: stage_impl %.1f ms\n", s, now_ms() - t_a);
            Bs.emplace_back(B);
            ChainSpec cs;
            if (s > 0) {
                cs.prev = Bs[(size_t)s - 1].get();
                cs.nIndices = nIndices;
                cs.L = L;
                cs.M = M;
                cs.pad_unif = pad_unif[s - 1];
                rc = chain_check(cs, B);
                if (rc != QUILT_OK) break;
            }
            const double t_b = now_ms();
            rc = pipe.push_batch(B, out[s], s > 0 ? &cs : nullptr);
            if (trace_on()) std::fprintf(stderr, "[quilt trace] chain stage %d: push_batch %.1f ms (fill %.1f, wait+unpack %.1f)\n", s, now_ms() - t_b, pipe.t_fill, pipe.t_unpack);
            pipe.t_fill = pipe.t_unpack = 0;
        }
        const double t_c = now_ms();
        if (rc == QUILT_OK) rc = pipe.drain();
        sync_all_streams();
        if (trace_on()) std::fprintf(stderr, "[quilt trace] chain drain %.1f ms\n", now_ms() - t_c);
    }
    const double t_d = now_ms();
    Bs.clear();  // (buffers go back to the caches)
    if (trace_on()) std::fprintf(stderr, "[quilt trace] chain teardown %.1f ms\n", now_ms() - t_d);
    return rc;
}

int quilt_gpu_gibbs(const QuiltGibbsArgs* args, QuiltGibbsOut* out) { return quilt_gpu_gibbs_batch(1, args, out); }

// ------------------------------------------------------------------------------------------------ haplotype re-selection
namespace {

struct SelPlan {
    SelParams P;
    size_t per_job = 0;   // scratch bytes per job
    size_t o_sym, o_runlen, o_rows, o_nrows, o_skey, o_sval, o_list, o_nlist, o_firstpos, o_tmp, o_ranked, o_tmp2, o_stage, o_pool, o_counts;
};

int make_sel_plan(const PanelDev& pd, int nHap, int Knew, int nIndices, int L, int M, int pad, SelPlan* pl) {
    if (nHap < 1 || nHap > 3 || Knew < 1 || nIndices < 1 || L < 1 || M < 1) return set_err(QUILT_ERR_BAD_ARG, "bad selection arguments");
    if (2 * L > SEL_MAXTOP) return set_err(QUILT_ERR_UNSUPPORTED, "mspbwtL > 8 not supported");
    if (pd.Tc < nIndices) return set_err(QUILT_ERR_BAD_ARG, "fewer grids than mspbwt_nindices (the reference resets nindices to 1, quilt-prepare-reference.R:452-455)");
    if (pd.K_full >= (1 << 24) || pd.Tc >= (1 << 20)) return set_err(QUILT_ERR_UNSUPPORTED, "panel too large for the packed match rows");
    SelParams& P = pl->P;
    std::memset(&P, 0, sizeof(P));
    P.K_full = pd.K_full;
    P.Tc = pd.Tc;
    P.nSNPs = pd.nSNPsC;
    P.nMaxDH = pd.nMaxDH;
    P.nHap = nHap;
    P.nIndices = nIndices;
    P.L = L;
    P.M = M;
    P.Knew = Knew;
    P.pad = pad;
    const int n_pos_max = (pd.Tc + nIndices - 1) / nIndices;
    P.rows_cap = n_pos_max * 2 * L;
    int sc = 1;
    while (sc < nIndices * P.rows_cap) sc <<= 1;
    P.sort_cap = sc;
    size_t o = 0;
    auto take = [&](size_t n) {
        const size_t r = o;
        o += al(n);
        return r;
    };
    const size_t hs = (size_t)nHap * sc;
    pl->o_sym = take((size_t)nHap * pd.Tc * 4);
    pl->o_runlen = take((size_t)nHap * nIndices * pd.K_full * 2);
    pl->o_rows = take((size_t)nHap * nIndices * P.rows_cap * 8);
    pl->o_nrows = take((size_t)nHap * nIndices * 4);
    pl->o_skey = take((size_t)sc * 8);
    pl->o_sval = take((size_t)sc * 4);
    pl->o_list = take(hs * 8);
    pl->o_nlist = take((size_t)nHap * 4);
    pl->o_firstpos = take(((size_t)std::max(pd.K_full, pd.Tc) + 2) * 4);
    pl->o_tmp = take(std::max(hs, (size_t)pd.K_full + 2) * 4);
    pl->o_ranked = take(hs * 4);
    pl->o_tmp2 = take(hs * 4);
    pl->o_stage = take((size_t)sc * 8);
    pl->o_pool = take(((size_t)pd.K_full + 2) * 4);
    pl->o_counts = take(16);
    pl->per_job = o;
    return QUILT_OK;
}

void fill_sel_job(const SelPlan& pl, char* scratch, const double* hapProbs, int32_t* which_out, const double* pad_unif, SelJob* J) {
    J->hapProbs = hapProbs;
    J->which_out = which_out;
    J->pad_unif = pad_unif;
    J->sym = (int32_t*)(scratch + pl.o_sym);
    J->runlen = (uint16_t*)(scratch + pl.o_runlen);
    J->rows = (uint64_t*)(scratch + pl.o_rows);
    J->nrows = (int32_t*)(scratch + pl.o_nrows);
    J->skey = (uint64_t*)(scratch + pl.o_skey);
    J->sval = (uint32_t*)(scratch + pl.o_sval);
    J->list = (uint64_t*)(scratch + pl.o_list);
    J->nlist = (int32_t*)(scratch + pl.o_nlist);
    J->firstpos = (int32_t*)(scratch + pl.o_firstpos);
    J->tmp = (int32_t*)(scratch + pl.o_tmp);
    J->ranked = (int32_t*)(scratch + pl.o_ranked);
    J->tmp2 = (int32_t*)(scratch + pl.o_tmp2);
    J->stage = (uint64_t*)(scratch + pl.o_stage);
    J->pool = (int32_t*)(scratch + pl.o_pool);
    J->counts_out = (int32_t*)(scratch + pl.o_counts);
}

int launch_select(const SelPlan& pl, const PanelDev& pd, const SelJob* djobs, int n) {
    const SelParams& P = pl.P;
    k_sel_symbols<<<dim3(P.Tc, P.nHap, n), SEL_NT, 0, g_stream>>>(P, djobs, pd.distinctHapsB, pd.n_used);
    LAUNCHED_K("k_sel_symbols");
    k_sel_match<<<dim3(P.nIndices, P.nHap, n), SEL_NT, 0, g_stream>>>(P, djobs, pd.hapMatcherR);
    LAUNCHED_K("k_sel_match");
    k_sel_rank<<<n, SEL_RANK_NT, 0, g_stream>>>(P, djobs);
    LAUNCHED_K("k_sel_rank");
    CK(cudaGetLastError());
    return QUILT_OK;
}

}  // namespace

int quilt_gpu_select_haps(const QuiltSelectArgs* a, int32_t* which_haps_to_use, int32_t* n_found, int32_t* n_unique) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!a || !a->panel || !a->hapProbs_t || !which_haps_to_use || !n_found || !n_unique) return set_err(QUILT_ERR_BAD_ARG, "null argument");
    int rc = ensure_device();
    if (rc != QUILT_OK) return rc;
    PanelDev pd;
    std::shared_ptr<PanelEntry> keep;
    if ((rc = get_panel(a->panel, &pd, &keep)) != QUILT_OK) return rc;
    SelPlan pl;
    if ((rc = make_sel_plan(pd, a->nHap, a->Knew, a->mspbwt_nindices, a->mspbwtL, a->mspbwtM, 0, &pl)) != QUILT_OK) return rc;
    const size_t b_hp = al((size_t)3 * pd.nSNPsC * 8), b_w = al((size_t)a->Knew * 4);
    DBuf buf;
    CK(buf.alloc(pl.per_job + b_hp + b_w + al(sizeof(SelJob))));
    char* d = (char*)buf.p;
    CK(cudaMemcpyAsync(d + pl.per_job, a->hapProbs_t, (size_t)3 * pd.nSNPsC * 8, cudaMemcpyHostToDevice, g_stream));
    SelJob J;
    fill_sel_job(pl, d, (const double*)(d + pl.per_job), (int32_t*)(d + pl.per_job + b_hp), nullptr, &J);
    SelJob* dj = (SelJob*)(d + pl.per_job + b_hp + b_w);
    CK(cudaMemcpyAsync(dj, &J, sizeof(SelJob), cudaMemcpyHostToDevice, g_stream));
    if ((rc = launch_select(pl, pd, dj, 1)) != QUILT_OK) return rc;
    int32_t counts[2] = {0, 0};
    CK(cudaMemcpyAsync(which_haps_to_use, J.which_out, (size_t)a->Knew * 4, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaMemcpyAsync(counts, J.counts_out, 8, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    *n_found = counts[0];
    *n_unique = counts[1];
    return QUILT_OK;
}

namespace {

DBuf g_sel_scratch;  // scratch of the chained selection (grown on demand; kernels on the library stream are serialised)

int chain_check(const ChainSpec& cs, QuiltGpuBatch* next) {
    QuiltGpuBatch* prev = cs.prev;
    if (!prev || !next) return set_err(QUILT_ERR_BAD_ARG, "null batch");
    if (!prev->ran) return set_err(QUILT_ERR_BAD_ARG, "the previous batch has not been run");
    if (prev->n != next->n) return set_err(QUILT_ERR_BAD_ARG, "chained batches must hold the same number of calls");
    if (prev->panel_ref != next->panel_ref) return set_err(QUILT_ERR_BAD_ARG, "chained batches must share one panel");
    const int K = next->jobs[0].a.K;
    const int nHap = (prev->jobs[0].a.flags & QUILT_F_SAMPLE_IS_DIPLOID) ? 2 : 3;
    for (int i = 0; i < prev->n; i++) {
        if (next->jobs[(size_t)i].a.K != K) return set_err(QUILT_ERR_UNSUPPORTED, "chained calls must share one Ksubset");
        if (prev->jobs[(size_t)i].a.nSNPs != prev->panel.nSNPsC) return set_err(QUILT_ERR_BAD_ARG, "selection runs on common-SNP calls");
        if (((prev->jobs[(size_t)i].a.flags & QUILT_F_SAMPLE_IS_DIPLOID) ? 2 : 3) != nHap) return set_err(QUILT_ERR_UNSUPPORTED, "mixed ploidy in a chained batch");
    }
    return QUILT_OK;
}

int chain_select_subset(const ChainSpec& cs, QuiltGpuBatch* next, const int* ids, int m) {
    QuiltGpuBatch* prev = cs.prev;
    const int K = next->jobs[0].a.K;
    const int nHap = (prev->jobs[0].a.flags & QUILT_F_SAMPLE_IS_DIPLOID) ? 2 : 3;
    SelPlan pl;
    int rc = make_sel_plan(prev->panel, nHap, K, cs.nIndices, cs.L, cs.M, 1, &pl);
    if (rc != QUILT_OK) return rc;
    // groups keep the scratch bounded
    const int group = std::max(1, std::min(m, (int)(((size_t)2 << 30) / pl.per_job)));
    const size_t b_pu = al((size_t)group * K * 8), b_j = al((size_t)group * sizeof(SelJob));
    const size_t need = (size_t)group * pl.per_job + b_pu + b_j;
    if (g_sel_scratch.bytes < need) {
        CK(cudaStreamSynchronize(g_stream));
        CK(g_sel_scratch.alloc(need));
    }
    char* d = (char*)g_sel_scratch.p;
    double* d_pu = (double*)(d + (size_t)group * pl.per_job);
    SelJob* dj = (SelJob*)(d + (size_t)group * pl.per_job + b_pu);
    std::vector<SelJob> hj((size_t)group);
    for (int j0 = 0; j0 < m; j0 += group) {
        const int mm = std::min(group, m - j0);
        if (j0 > 0) CK(cudaStreamSynchronize(g_stream));  // the previous group still reads the scratch
        for (int q = 0; q < mm; q++) {
            const int id = ids[j0 + q];
            const HostJob& pj = prev->jobs[(size_t)id];
            const HostJob& nj = next->jobs[(size_t)id];
            const double* hp = pj.lo.hap_dev_only ? (const double*)((const char*)prev->doutB.p + pj.outB_off + pj.lo.hap)
                                                  : (const double*)((const char*)prev->dout().p + pj.out_off + pj.lo.hap);
            int32_t* which = (int32_t*)((char*)next->din().p + nj.in_off + nj.li.which);
            CK(cudaMemcpyAsync(d_pu + (size_t)q * K, cs.pad_unif + (size_t)id * K, (size_t)K * 8, cudaMemcpyHostToDevice, g_stream));
            fill_sel_job(pl, d + (size_t)q * pl.per_job, hp, which, d_pu + (size_t)q * K, &hj[(size_t)q]);
        }
        CK(cudaMemcpyAsync(dj, hj.data(), (size_t)mm * sizeof(SelJob), cudaMemcpyHostToDevice, g_stream));
        if ((rc = launch_select(pl, prev->panel, dj, mm)) != QUILT_OK) return rc;
    }
    return QUILT_OK;
}

}  // namespace

int quilt_gpu_batch_chain_select(QuiltGpuBatch* prev, QuiltGpuBatch* next, int32_t nIndices, int32_t L, int32_t M, const double* pad_unif) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!prev || !next || !pad_unif) return set_err(QUILT_ERR_BAD_ARG, "null argument");
    ChainSpec cs;
    cs.prev = prev;
    cs.nIndices = nIndices;
    cs.L = L;
    cs.M = M;
    cs.pad_unif = pad_unif;
    int rc = chain_check(cs, next);
    if (rc != QUILT_OK) return rc;
    std::vector<int> ids((size_t)prev->n);
    for (int i = 0; i < prev->n; i++) ids[(size_t)i] = i;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, g_stream));
    if ((rc = chain_select_subset(cs, next, ids.data(), prev->n)) != QUILT_OK) return rc;
    CK(cudaEventRecord(e1, g_stream));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    next->chain_ms = ms;
    next->chained = true;
    return QUILT_OK;
}

int quilt_gpu_batch_chain_timing(QuiltGpuBatch* B, double* chain_ms) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!B || !chain_ms) return set_err(QUILT_ERR_BAD_ARG, "null argument");
    *chain_ms = B->chain_ms;
    return QUILT_OK;
}

int quilt_gpu_batch_which_haps(QuiltGpuBatch* B, int32_t job, int32_t* which_haps_to_use) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!B || job < 0 || job >= B->n || !which_haps_to_use) return set_err(QUILT_ERR_BAD_ARG, "bad argument");
    const HostJob& j = B->jobs[(size_t)job];
    CK(cudaMemcpyAsync(which_haps_to_use, (const char*)B->din().p + j.in_off + j.li.which, (size_t)j.a.K * 4, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    return QUILT_OK;
}




// ------------------------------------------------------------------------------------------------ full-panel haploid pass
namespace {
double g_hap_ms = 0, g_hap_bytes = 0;

int launch_hap_fb(int ept, const HapParams& P, const HapJob* dj, int n, const PanelDev& pd) {
    const size_t sm = hap_smem_bytes(P.nMaxDH, P.K);
    if (sm > 227 * 1024) return set_err(QUILT_ERR_UNSUPPORTED, "full-panel pass: the panel does not fit one CTA's shared memory");
#define QB_HAP(E)                                                                                      \
    {                                                                                                  \
        CK(cudaFuncSetAttribute(k_hap_fb<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));   \
        k_hap_fb<E><<<n, HAP_NT, sm, g_stream>>>(P, dj, pd);                                           \
    }
    if (ept <= 4) QB_HAP(4)
    else if (ept <= 10) QB_HAP(10)
    else if (ept <= 16) QB_HAP(16)
    else QB_HAP(32)
#undef QB_HAP
    LAUNCHED_K("k_hap_fb");
    return QUILT_OK;
}
}  // namespace

namespace {
// haplotypes of every grid sorted by symbol (stable counting sort), once per panel
int ensure_hap_index(const QuiltPanel* p, PanelEntry& e) {
    if (e.hap_index.p) {
        e.dev.hap_perm = (const uint16_t*)e.hap_index.p;
        e.dev.hap_symoff = (const int32_t*)((const char*)e.hap_index.p + al((size_t)p->K_full * p->nGrids * 2));
        return QUILT_OK;
    }
    const int K = p->K_full, T = p->nGrids, NM2 = p->nMaxDH + 2;
    if (K > 65536) return set_err(QUILT_ERR_UNSUPPORTED, "full-panel pass: K_full > 65536");
    std::vector<uint16_t> perm((size_t)K * T);
    std::vector<int32_t> off((size_t)T * NM2);
    parallel_for(T, [&](int g) {
        const uint8_t* col = p->hapMatcherR + (size_t)g * K;
        int32_t* o = off.data() + (size_t)g * NM2;
        std::vector<int32_t> cnt((size_t)NM2, 0);
        for (int k = 0; k < K; k++) cnt[std::min<int>(col[k], p->nMaxDH) + 1]++;
        for (int i = 1; i < NM2; i++) cnt[(size_t)i] += cnt[(size_t)i - 1];
        for (int i = 0; i < NM2; i++) o[i] = cnt[(size_t)i];
        std::vector<int32_t> pos(cnt.begin(), cnt.end() - 1);
        uint16_t* pg = perm.data() + (size_t)g * K;
        for (int k = 0; k < K; k++) pg[pos[std::min<int>(col[k], p->nMaxDH)]++] = (uint16_t)k;
    });
    const size_t b_perm = al(perm.size() * 2);
    CK(e.hap_index.alloc(b_perm + al(off.size() * 4)));
    CK(cudaMemcpy(e.hap_index.p, perm.data(), perm.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy((char*)e.hap_index.p + b_perm, off.data(), off.size() * 4, cudaMemcpyHostToDevice));
    e.dev.hap_perm = (const uint16_t*)e.hap_index.p;
    e.dev.hap_symoff = (const int32_t*)((const char*)e.hap_index.p + b_perm);
    return QUILT_OK;
}
}  // namespace

int quilt_gpu_haploid_dosage_versus_refs_batch(int32_t n, const QuiltHaploidArgs* args, QuiltHaploidOut* out) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (n < 1 || !args || !out || !args[0].panel) return set_err(QUILT_ERR_BAD_ARG, "bad arguments");
    int rc = ensure_device();
    if (rc != QUILT_OK) return rc;
    PanelDev pd;
    std::shared_ptr<PanelEntry> keep;
    if ((rc = get_panel(args[0].panel, &pd, &keep)) != QUILT_OK) return rc;
    if ((rc = ensure_hap_index(args[0].panel, *keep)) != QUILT_OK) return rc;
    pd = keep->dev;
    const int K = pd.K_full, T = pd.Tc, nS = pd.nSNPsC, NM1 = pd.nMaxDH + 1;
    if (T < 2) return set_err(QUILT_ERR_BAD_ARG, "the full-panel pass needs at least two grids");
    if (K > HAP_NT * 32) return set_err(QUILT_ERR_UNSUPPORTED, "full-panel pass: K_full > 16384 needs the multi-CTA form (not built); use the mspbwt selection for such panels");
    HapParams P;
    std::memset(&P, 0, sizeof(P));
    P.K = K;
    P.T = T;
    P.nSNPs = nS;
    P.nMaxDH = pd.nMaxDH;
    P.n_thin = args[0].n_thinned;
    P.K_top = args[0].K_top_matches;
    P.best_cap = args[0].best_cap;
    P.flags = args[0].flags;
    P.thr = args[0].min_emission_prob_normalization_threshold;
    P.ref_error = pd.ref_error;
    for (int i = 0; i < n; i++) {
        const QuiltHaploidArgs& a = args[i];
        if (!a.gl || !a.transMatRate_t) return set_err(QUILT_ERR_BAD_ARG, "gl / transMatRate_t is NULL");
        if (a.panel != args[0].panel && !same_panel_struct(*a.panel, *args[0].panel)) return set_err(QUILT_ERR_UNSUPPORTED, "all passes of a batch must share one panel");
        if (a.flags != P.flags || a.n_thinned != P.n_thin || a.K_top_matches != P.K_top || a.best_cap != P.best_cap ||
            a.min_emission_prob_normalization_threshold != P.thr)
            return set_err(QUILT_ERR_UNSUPPORTED, "all passes of a batch must share flags / thinning / K_top_matches");
        if ((a.flags & QUILT_HF_GET_BEST_HAPS) && (!a.gammaSmall_cols_to_get || a.n_thinned < 1 || a.best_cap < 1)) return set_err(QUILT_ERR_BAD_ARG, "best haplotypes need gammaSmall_cols_to_get, n_thinned and best_cap");
    }
    if (P.K_top < 1 || P.K_top > HAP_MAXTOP) return set_err(QUILT_ERR_UNSUPPORTED, "K_top_matches must be in 1 .. 16");
    const bool w_beta = (P.flags & QUILT_HF_RETURN_BETAHAT) != 0, w_gamma = (P.flags & QUILT_HF_RETURN_GAMMA) != 0, w_best = (P.flags & QUILT_HF_GET_BEST_HAPS) != 0;
    // per-pass device layout
    size_t o = 0;
    auto take = [&](size_t nb) {
        const size_t r = o;
        o += al(nb);
        return r;
    };
    const size_t KT = (size_t)K * T;
    const size_t o_gl = take((size_t)nS * 16), o_tm = take((size_t)(T - 1) * 16), o_cols = take((size_t)T * 4), o_em = take((size_t)T * NM1 * 8);
    const size_t o_emax = take((size_t)T * 8), o_cmin = take((size_t)T * 8), o_hv = take((size_t)T), o_alpha = take(KT * 8);
    const size_t o_beta = take(w_beta ? KT * 8 : 0), o_gamma = take(w_gamma ? KT * 8 : 0), o_c = take((size_t)T * 8), o_dos = take((size_t)nS * 8);
    const size_t nb_best = w_best ? (size_t)P.n_thin * P.best_cap : 0;
    const size_t o_best = take(nb_best * 4), o_bval = take(nb_best * 8), o_bcnt = take((size_t)std::max(P.n_thin, 1) * 4);
    const size_t per = o;
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    const int group = (int)std::max<size_t>(1, std::min<size_t>((size_t)n, (size_t)(free_b * 0.8) / per));
    // (the state of a group of passes is tens of GB: the buffers are recycled across calls like the Gibbs staging buffers)
    struct Recycled {
        DBuf b;
        ~Recycled() {
            if (b.bytes > ((size_t)16 << 30))
                b.release();  // (a very large group state is not worth keeping away from the Gibbs batches)
            else
                cached_release(b, g_dcache_hap);
        }
    } buf_r, jb_r;
    DBuf& buf = buf_r.b;
    DBuf& jb = jb_r.b;
    CK(cached_alloc(buf, g_dcache_hap, (size_t)group * per));
    CK(cached_alloc(jb, g_dcache_hap, (size_t)group * sizeof(HapJob)));
    std::vector<HapJob> hj((size_t)group);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    double ms_total = 0;
    for (int j0 = 0; j0 < n; j0 += group) {
        const int m = std::min(group, n - j0);
        for (int q = 0; q < m; q++) {
            const QuiltHaploidArgs& a = args[j0 + q];
            char* d = (char*)buf.p + (size_t)q * per;
            CK(cudaMemcpyAsync(d + o_gl, a.gl, (size_t)nS * 16, cudaMemcpyHostToDevice, g_stream));
            CK(cudaMemcpyAsync(d + o_tm, a.transMatRate_t, (size_t)(T - 1) * 16, cudaMemcpyHostToDevice, g_stream));
            if (w_best) CK(cudaMemcpyAsync(d + o_cols, a.gammaSmall_cols_to_get, (size_t)T * 4, cudaMemcpyHostToDevice, g_stream));
            HapJob& J = hj[(size_t)q];
            J.gl = (const double*)(d + o_gl);
            J.tm = (const double*)(d + o_tm);
            J.cols = w_best ? (const int32_t*)(d + o_cols) : nullptr;
            J.eMatDH = (double*)(d + o_em);
            J.emax = (double*)(d + o_emax);
            J.cmin = (double*)(d + o_cmin);
            J.hasvar = (uint8_t*)(d + o_hv);
            J.alpha = (double*)(d + o_alpha);
            J.beta = w_beta ? (double*)(d + o_beta) : nullptr;
            J.gamma = w_gamma ? (double*)(d + o_gamma) : nullptr;
            J.c = (double*)(d + o_c);
            J.dosage = (double*)(d + o_dos);
            J.best = (int32_t*)(d + o_best);
            J.best_val = (double*)(d + o_bval);
            J.best_cnt = (int32_t*)(d + o_bcnt);
        }
        CK(cudaMemcpyAsync(jb.p, hj.data(), (size_t)m * sizeof(HapJob), cudaMemcpyHostToDevice, g_stream));
        CK(cudaEventRecord(e0, g_stream));
        k_hap_eMatDH<<<dim3(T, m), 256, 0, g_stream>>>(P, (const HapJob*)jb.p, pd.distinctHapsB);
        LAUNCHED_K("k_hap_eMatDH");
        rc = launch_hap_fb((K + HAP_NT - 1) / HAP_NT, P, (const HapJob*)jb.p, m, pd);
        if (rc != QUILT_OK) return rc;
        CK(cudaEventRecord(e1, g_stream));
        CK(cudaGetLastError());
        for (int q = 0; q < m; q++) {
            QuiltHaploidOut& O = out[j0 + q];
            const char* d = (const char*)buf.p + (size_t)q * per;
            if (O.dosage && (P.flags & QUILT_HF_RETURN_DOSAGE)) CK(cudaMemcpyAsync(O.dosage, d + o_dos, (size_t)nS * 8, cudaMemcpyDeviceToHost, g_stream));
            if (O.c) CK(cudaMemcpyAsync(O.c, d + o_c, (size_t)T * 8, cudaMemcpyDeviceToHost, g_stream));
            if (O.alphaHat_t && (P.flags & QUILT_HF_RETURN_ALPHAHAT)) CK(cudaMemcpyAsync(O.alphaHat_t, d + o_alpha, KT * 8, cudaMemcpyDeviceToHost, g_stream));
            if (O.betaHat_t && w_beta) CK(cudaMemcpyAsync(O.betaHat_t, d + o_beta, KT * 8, cudaMemcpyDeviceToHost, g_stream));
            if (O.gamma_t && w_gamma) CK(cudaMemcpyAsync(O.gamma_t, d + o_gamma, KT * 8, cudaMemcpyDeviceToHost, g_stream));
            if (w_best) {
                if (O.best_haps) CK(cudaMemcpyAsync(O.best_haps, d + o_best, nb_best * 4, cudaMemcpyDeviceToHost, g_stream));
                if (O.best_haps_values) CK(cudaMemcpyAsync(O.best_haps_values, d + o_bval, nb_best * 8, cudaMemcpyDeviceToHost, g_stream));
                if (O.best_haps_count) CK(cudaMemcpyAsync(O.best_haps_count, d + o_bcnt, (size_t)P.n_thin * 4, cudaMemcpyDeviceToHost, g_stream));
            }
        }
        CK(cudaStreamSynchronize(g_stream));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        ms_total += ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    g_hap_ms = ms_total;
    g_hap_bytes = (double)n * (double)KT * (2.0 + 16.0 + (w_beta ? 8.0 : 0.0) + (w_gamma ? 8.0 : 0.0));
    return QUILT_OK;
}

int quilt_gpu_haploid_dosage_versus_refs(const QuiltHaploidArgs* args, QuiltHaploidOut* out) {
    return quilt_gpu_haploid_dosage_versus_refs_batch(1, args, out);
}

int quilt_gpu_haploid_last_timing(double* kernel_ms, double* algorithmic_bytes) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (kernel_ms) *kernel_ms = g_hap_ms;
    if (algorithmic_bytes) *algorithmic_bytes = g_hap_bytes;
    return QUILT_OK;
}


// ------------------------------------------------------------------------------------------------ per-sample summary
namespace {
const double* job_hap_dev(const QuiltGpuBatch* B, int job) {
    const HostJob& j = B->jobs[(size_t)job];
    return j.lo.hap_dev_only ? (const double*)((const char*)B->doutB.p + j.outB_off + j.lo.hap) : (const double*)((const char*)B->dout().p + j.out_off + j.lo.hap);
}
}  // namespace

int quilt_gpu_samples_summary(int32_t n_samples, int32_t nSNPs, const QuiltSampleSummary* S, double* infoCount, double* afCount, double* hweCount) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (n_samples < 1 || nSNPs < 1 || !S) return set_err(QUILT_ERR_BAD_ARG, "bad arguments");
    int rc = ensure_device();
    if (rc != QUILT_OK) return rc;
    size_t n_ptr = 0;
    for (int q = 0; q < n_samples; q++) {
        const QuiltSampleSummary& s = S[q];
        if (s.n_calls < 1 || !s.calls || !s.phasing.batch) return set_err(QUILT_ERR_BAD_ARG, "sample without stored calls / phasing call");
        for (int c = 0; c <= s.n_calls; c++) {
            const QuiltSummaryCall& cc = c < s.n_calls ? s.calls[c] : s.phasing;
            if (!cc.batch || !cc.batch->ran || cc.job < 0 || cc.job >= cc.batch->n) return set_err(QUILT_ERR_BAD_ARG, "summary call refers to a batch that has not run / a bad job");
            const QuiltGibbsArgs& a = cc.batch->jobs[(size_t)cc.job].a;
            if (a.nSNPs != nSNPs) return set_err(QUILT_ERR_BAD_ARG, "summary calls must all be on the same SNP axis");
            if (!(a.flags & QUILT_F_SAMPLE_IS_DIPLOID)) return set_err(QUILT_ERR_UNSUPPORTED, "the device summary handles diploid samples (recast_nipt_haps is not built)");
        }
        n_ptr += (size_t)s.n_calls;
    }
    const size_t ns = (size_t)nSNPs;
    const size_t per = al(ns * 8) + al(ns * 24) + al(ns * 16) + al(ns * 2) + 2 * al(ns * 8) + al(ns);
    const size_t b_ptr = al(n_ptr * 8), b_s = al((size_t)n_samples * sizeof(SumSample)), b_cnt = al(ns * 16) + al(ns * 8) + al(ns * 24);
    DBuf buf;
    CK(buf.alloc((size_t)n_samples * per + b_ptr + b_s + b_cnt));
    char* d = (char*)buf.p;
    const double** d_ptr = (const double**)(d + (size_t)n_samples * per);
    SumSample* d_s = (SumSample*)(d + (size_t)n_samples * per + b_ptr);
    char* d_cnt = d + (size_t)n_samples * per + b_ptr + b_s;
    std::vector<const double*> h_ptr;
    std::vector<SumSample> h_s((size_t)n_samples);
    for (int q = 0; q < n_samples; q++) {
        const QuiltSampleSummary& s = S[q];
        SumSample& J = h_s[(size_t)q];
        J.call_hp = d_ptr + h_ptr.size();
        for (int c = 0; c < s.n_calls; c++) h_ptr.push_back(job_hap_dev(s.calls[c].batch, s.calls[c].job));
        J.phase_hp = job_hap_dev(s.phasing.batch, s.phasing.job);
        J.n_calls = s.n_calls;
        char* o = d + (size_t)q * per;
        J.dosage = (double*)o, o += al(ns * 8);
        J.gp = (double*)o, o += al(ns * 24);
        J.hd = (double*)o, o += al(ns * 16);
        J.gt = (int8_t*)o, o += al(ns * 2);
        J.eij = (double*)o, o += al(ns * 8);
        J.fij = (double*)o, o += al(ns * 8);
        J.maxgen = (int8_t*)o;
    }
    CK(cudaMemcpyAsync(d_ptr, h_ptr.data(), n_ptr * 8, cudaMemcpyHostToDevice, g_stream));
    CK(cudaMemcpyAsync(d_s, h_s.data(), (size_t)n_samples * sizeof(SumSample), cudaMemcpyHostToDevice, g_stream));
    k_sample_summary<<<dim3((nSNPs + 255) / 256, n_samples), 256, 0, g_stream>>>(d_s, nSNPs);
    LAUNCHED_K("k_sample_summary");
    double* d_info = (double*)d_cnt;
    double* d_af = (double*)(d_cnt + al(ns * 16));
    double* d_hwe = (double*)(d_cnt + al(ns * 16) + al(ns * 8));
    k_info_counts<<<(nSNPs + 255) / 256, 256, 0, g_stream>>>(d_s, n_samples, nSNPs, d_info, d_af, d_hwe);
    LAUNCHED_K("k_info_counts");
    CK(cudaGetLastError());
    for (int q = 0; q < n_samples; q++) {
        const QuiltSampleSummary& s = S[q];
        const SumSample& J = h_s[(size_t)q];
        if (s.dosage) CK(cudaMemcpyAsync(s.dosage, J.dosage, ns * 8, cudaMemcpyDeviceToHost, g_stream));
        if (s.gp_t) CK(cudaMemcpyAsync(s.gp_t, J.gp, ns * 24, cudaMemcpyDeviceToHost, g_stream));
        if (s.hd) CK(cudaMemcpyAsync(s.hd, J.hd, ns * 16, cudaMemcpyDeviceToHost, g_stream));
        if (s.gt) CK(cudaMemcpyAsync(s.gt, J.gt, ns * 2, cudaMemcpyDeviceToHost, g_stream));
    }
    if (infoCount) CK(cudaMemcpyAsync(infoCount, d_info, ns * 16, cudaMemcpyDeviceToHost, g_stream));
    if (afCount) CK(cudaMemcpyAsync(afCount, d_af, ns * 8, cudaMemcpyDeviceToHost, g_stream));
    if (hweCount) CK(cudaMemcpyAsync(hweCount, d_hwe, ns * 24, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    return QUILT_OK;
}

// ------------------------------------------------------------------------------------------------ ingestion / VCF column
int quilt_gpu_ingest_pileup(const QuiltPileup* in, QuiltIngestOut* out) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!in || !out || in->nReads < 1 || in->nSNPs < 1 || in->nGrids < 1 || !in->offsets || !in->u || !in->bq || !in->central_snp || !in->grid)
        return set_err(QUILT_ERR_BAD_ARG, "bad pileup");
    int rc = ensure_device();
    if (rc != QUILT_OK) return rc;
    const int R = in->nReads, nS = in->nSNPs, T = in->nGrids, nU = in->offsets[R];
    if (nU < R) return set_err(QUILT_ERR_BAD_ARG, "every read needs at least one SNP");
    for (int r = 0; r < R; r++)
        if (in->central_snp[r] < 0 || in->central_snp[r] >= nS) return set_err(QUILT_ERR_BAD_ARG, "central SNP outside [0, nSNPs)");
    for (int s = 0; s < nS; s++)
        if (in->grid[s] < 0 || in->grid[s] >= T) return set_err(QUILT_ERR_BAD_ARG, "grid outside [0, nGrids)");
    // convertScaledBQtoProbs (STITCH): bq < 0 -> (1 - eps, eps / 3), bq > 0 -> (eps / 3, 1 - eps) with eps = 10^(-|bq| / 10), libm pow
    std::vector<double> prob((size_t)nU * 2, 0.0);
    for (int t = 0; t < nU; t++) {
        if (in->u[t] < 0 || in->u[t] >= nS) return set_err(QUILT_ERR_BAD_ARG, "SNP index outside [0, nSNPs)");
        const int bq = in->bq[t];
        if (bq < 0) {
            const double eps = std::pow(10, double(bq) / 10);
            prob[2 * (size_t)t] = 1 - eps;
            prob[2 * (size_t)t + 1] = eps * (1.0 / 3.0);
        } else if (bq > 0) {
            const double eps = std::pow(10, -double(bq) / 10);
            prob[2 * (size_t)t] = eps * (1.0 / 3.0);
            prob[2 * (size_t)t + 1] = 1 - eps;
        }
    }
    size_t o = 0;
    auto take = [&](size_t nb) {
        const size_t r = o;
        o += al(nb);
        return r;
    };
    const size_t o_off = take((size_t)(R + 1) * 4), o_u = take((size_t)nU * 4), o_bq = take((size_t)nU * 4), o_cen = take((size_t)R * 4), o_grid = take((size_t)nS * 4);
    const size_t o_prob = take((size_t)nU * 16), o_wif = take((size_t)R * 4);
    const size_t o_zero = o;  // counters (zeroed)
    const size_t o_cntG = take((size_t)(T + 1) * 4), o_fillG = take((size_t)T * 4), o_cntS = take((size_t)(nS + 1) * 4), o_fillS = take((size_t)nS * 4);
    const size_t o_zero_end = o;
    const size_t o_ordG = take((size_t)R * 4), o_ordS = take((size_t)nU * 4), o_ncnt = take((size_t)(R + 1) * 4);
    const size_t o_us = take((size_t)nU * 4), o_bqs = take((size_t)nU * 4), o_wifs = take((size_t)R * 4), o_ac = take((size_t)nS * 16), o_ghr = take((size_t)T);
    DBuf buf;
    CK(buf.alloc(o));
    char* d = (char*)buf.p;
    CK(cudaMemcpyAsync(d + o_off, in->offsets, (size_t)(R + 1) * 4, cudaMemcpyHostToDevice, g_stream));
    CK(cudaMemcpyAsync(d + o_u, in->u, (size_t)nU * 4, cudaMemcpyHostToDevice, g_stream));
    CK(cudaMemcpyAsync(d + o_bq, in->bq, (size_t)nU * 4, cudaMemcpyHostToDevice, g_stream));
    CK(cudaMemcpyAsync(d + o_cen, in->central_snp, (size_t)R * 4, cudaMemcpyHostToDevice, g_stream));
    CK(cudaMemcpyAsync(d + o_grid, in->grid, (size_t)nS * 4, cudaMemcpyHostToDevice, g_stream));
    CK(cudaMemcpyAsync(d + o_prob, prob.data(), (size_t)nU * 16, cudaMemcpyHostToDevice, g_stream));
    CK(cudaMemsetAsync(d + o_zero, 0, o_zero_end - o_zero, g_stream));
    IngestDev D;
    D.R = R, D.nU = nU, D.nSNPs = nS, D.T = T;
    D.off = (const int32_t*)(d + o_off), D.u = (const int32_t*)(d + o_u), D.bq = (const int32_t*)(d + o_bq), D.central = (const int32_t*)(d + o_cen);
    D.grid = (const int32_t*)(d + o_grid), D.prob = (const double*)(d + o_prob), D.wif = (int32_t*)(d + o_wif);
    D.cntG = (int32_t*)(d + o_cntG), D.fillG = (int32_t*)(d + o_fillG), D.ordG = (int32_t*)(d + o_ordG);
    D.cntS = (int32_t*)(d + o_cntS), D.fillS = (int32_t*)(d + o_fillS), D.ordS = (int32_t*)(d + o_ordS), D.ncnt = (int32_t*)(d + o_ncnt);
    D.u_s = (int32_t*)(d + o_us), D.bq_s = (int32_t*)(d + o_bqs), D.wif_s = (int32_t*)(d + o_wifs), D.alleleCount = (double*)(d + o_ac), D.grid_has_read = (uint8_t*)(d + o_ghr);
    const int nmax = std::max(std::max(R, nU), std::max(T, nS));
    k_ing_count<<<(nmax + 255) / 256, 256, 0, g_stream>>>(D);
    LAUNCHED_K("k_ing_count");
    k_exscan_i32<<<1, 1024, 0, g_stream>>>(D.cntG, T);
    LAUNCHED_K("k_exscan_i32");
    k_exscan_i32<<<1, 1024, 0, g_stream>>>(D.cntS, nS);
    LAUNCHED_K("k_exscan_i32");
    k_ing_place<<<(nmax + 255) / 256, 256, 0, g_stream>>>(D);
    LAUNCHED_K("k_ing_place");
    k_ing_sort<<<(nmax + 255) / 256, 256, 0, g_stream>>>(D);
    LAUNCHED_K("k_ing_sort");
    k_ing_lens<<<(R + 255) / 256, 256, 0, g_stream>>>(D);
    LAUNCHED_K("k_ing_lens");
    k_exscan_i32<<<1, 1024, 0, g_stream>>>(D.ncnt, R);
    LAUNCHED_K("k_exscan_i32");
    k_ing_copy<<<R, 256, 0, g_stream>>>(D);
    LAUNCHED_K("k_ing_copy");
    k_ing_allele<<<(nS + 255) / 256, 256, 0, g_stream>>>(D);
    LAUNCHED_K("k_ing_allele");
    CK(cudaGetLastError());
    if (out->order) CK(cudaMemcpyAsync(out->order, D.ordG, (size_t)R * 4, cudaMemcpyDeviceToHost, g_stream));
    if (out->offsets) CK(cudaMemcpyAsync(out->offsets, D.ncnt, (size_t)(R + 1) * 4, cudaMemcpyDeviceToHost, g_stream));
    if (out->u) CK(cudaMemcpyAsync(out->u, D.u_s, (size_t)nU * 4, cudaMemcpyDeviceToHost, g_stream));
    if (out->bq) CK(cudaMemcpyAsync(out->bq, D.bq_s, (size_t)nU * 4, cudaMemcpyDeviceToHost, g_stream));
    if (out->wif0) CK(cudaMemcpyAsync(out->wif0, D.wif_s, (size_t)R * 4, cudaMemcpyDeviceToHost, g_stream));
    if (out->first_read_of_grid) CK(cudaMemcpyAsync(out->first_read_of_grid, D.cntG, (size_t)(T + 1) * 4, cudaMemcpyDeviceToHost, g_stream));
    if (out->grid_has_read) CK(cudaMemcpyAsync(out->grid_has_read, D.grid_has_read, (size_t)T, cudaMemcpyDeviceToHost, g_stream));
    if (out->alleleCount) CK(cudaMemcpyAsync(out->alleleCount, D.alleleCount, (size_t)nS * 16, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    return QUILT_OK;
}

int quilt_gpu_make_vcf_column(int32_t nSNPs, const double* gp_t, const double* hd, char* out) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (nSNPs < 1 || !gp_t || !hd || !out) return set_err(QUILT_ERR_BAD_ARG, "bad arguments");
    int rc = ensure_device();
    if (rc != QUILT_OK) return rc;
    const size_t ns = (size_t)nSNPs;
    DBuf buf;
    CK(buf.alloc(al(ns * 24) + al(ns * 16) + al(ns * VCF_REC)));
    char* d = (char*)buf.p;
    double* d_gp = (double*)d;
    double* d_hd = (double*)(d + al(ns * 24));
    char* d_out = d + al(ns * 24) + al(ns * 16);
    CK(cudaMemcpyAsync(d_gp, gp_t, ns * 24, cudaMemcpyHostToDevice, g_stream));
    CK(cudaMemcpyAsync(d_hd, hd, ns * 16, cudaMemcpyHostToDevice, g_stream));
    k_vcf_column<<<(nSNPs + 255) / 256, 256, 0, g_stream>>>(nSNPs, d_gp, d_hd, d_out);
    LAUNCHED_K("k_vcf_column");
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, d_out, ns * VCF_REC, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    return QUILT_OK;
}

// ---- component entry points (parity tests of the individual reference functions)
int quilt_gpu_make_eMatRead_t(const QuiltGibbsArgs* args, double* eMatRead_t, int32_t* read_category) {
    if (!args || !eMatRead_t) return set_err(QUILT_ERR_BAD_ARG, "null argument");
    QuiltGibbsArgs a = *args;
    a.flags |= QUILT_F_RETURN_EXTRA;
    QuiltGpuBatch* B = nullptr;
    int rc = quilt_gpu_batch_stage(1, &a, &B);
    if (rc != QUILT_OK) return rc;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        rc = cudaMemsetAsync(B->dout().p, 0, B->out_bytes, g_stream) == cudaSuccess ? QUILT_OK : set_err(QUILT_ERR_CUDA, "memset");
        if (rc == QUILT_OK && B->outB_bytes) rc = cudaMemsetAsync(B->doutB.p, 0, B->outB_bytes, g_stream) == cudaSuccess ? QUILT_OK : set_err(QUILT_ERR_CUDA, "memset");
        if (rc == QUILT_OK) rc = ensure_slots(B);
        if (rc == QUILT_OK) rc = run_bucket(B, *B->buckets[0], false, true);
        if (rc == QUILT_OK) {
            const HostJob& j = B->jobs[0];
            std::memcpy(eMatRead_t, j.dbg_eMatRead.data(), j.dbg_eMatRead.size() * 8);
            if (read_category) {
                cudaError_t e = cudaMemcpy(read_category, (const char*)B->dout().p + j.out_off + j.lo.cat, (size_t)j.R * 4, cudaMemcpyDeviceToHost);
                if (e != cudaSuccess) rc = set_err(QUILT_ERR_CUDA, cudaGetErrorString(e));
            }
        }
    }
    quilt_gpu_batch_free(B);
    return rc;
}

int quilt_gpu_unpack_panel(const QuiltPanel* panel, int32_t K, const int32_t* which, int32_t all_snps, uint32_t* words) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!panel || !which || !words || K <= 0) return set_err(QUILT_ERR_BAD_ARG, "null argument");
    int rc = ensure_device();
    if (rc != QUILT_OK) return rc;
    PanelDev PD;
    if ((rc = get_panel(panel, &PD)) != QUILT_OK) return rc;
    if (all_snps && panel->nSNPs_all <= 0) return set_err(QUILT_ERR_BAD_ARG, "panel has no all-SNP axis");
    for (int k = 0; k < K; k++)
        if (which[k] < 1 || which[k] > panel->K_full) return set_err(QUILT_ERR_BAD_ARG, "which_haps_to_use out of range");
    const int Kp = (K + 31) & ~31;
    const int T = all_snps ? (panel->nSNPs_all + 31) / 32 : panel->nGrids;
    DBuf dw, dwc, dwhich, dj, dtype;
    CK(dw.alloc((size_t)T * Kp * 4));
    CK(dwc.alloc((size_t)panel->nGrids * Kp * 4));
    CK(dwhich.alloc((size_t)K * 4));
    CK(dj.alloc(sizeof(JobDev)));
    CK(dtype.alloc(std::max(panel->nSNPs_all, 1)));
    CK(cudaMemcpy(dwhich.p, which, (size_t)K * 4, cudaMemcpyHostToDevice));
    JobDev J;
    std::memset(&J, 0, sizeof(J));
    J.W = (uint32_t*)dw.p;
    J.Wc = (uint32_t*)dwc.p;
    J.which = (const int32_t*)dwhich.p;
    J.snp_type = (uint8_t*)dtype.p;
    CK(cudaMemcpy(dj.p, &J, sizeof(J), cudaMemcpyHostToDevice));
    const JobDev* d = (const JobDev*)dj.p;
    const int kb = (Kp + 255) / 256;
    if (all_snps) {
        k_unpack_common<<<dim3(kb, PD.Tc, 1), 256, 0, g_stream>>>(PD, d, K, Kp, 1);
        LAUNCHED_K("k_unpack_common");
        k_assemble_all<<<dim3(kb, (T + ASM_GPB - 1) / ASM_GPB, 1), 256, 0, g_stream>>>(PD, d, K, Kp, T);
        LAUNCHED_K("k_assemble_all");
        k_scatter_rare<<<dim3((K + 255) / 256, 1), 256, 0, g_stream>>>(PD, d, K, Kp);
        LAUNCHED_K("k_scatter_rare");
    } else {
        k_unpack_common<<<dim3(kb, T, 1), 256, 0, g_stream>>>(PD, d, K, Kp, 0);
        LAUNCHED_K("k_unpack_common");
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(g_stream));
    std::vector<uint32_t> tmp((size_t)T * Kp);
    CK(cudaMemcpy(tmp.data(), dw.p, (size_t)T * Kp * 4, cudaMemcpyDeviceToHost));
    for (int g = 0; g < T; g++) std::memcpy(words + (size_t)g * K, &tmp[(size_t)g * Kp], (size_t)K * 4);
    return QUILT_OK;
}

int quilt_gpu_forward_backward(int32_t K, int32_t T, const double* eMatGrid_t, const double* tm, double* alphaHat_t, double* betaHat_t, double* c) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (K <= 0 || T <= 0 || !eMatGrid_t || !tm || !alphaHat_t || !betaHat_t || !c) return set_err(QUILT_ERR_BAD_ARG, "null argument");
    int rc = ensure_device();
    if (rc != QUILT_OK) return rc;
    Geo geo;
    if (!pick_geo(K, &geo)) return set_err(QUILT_ERR_UNSUPPORTED, "K > 8192 not supported");
    const int Kp = (K + 31) & ~31;
    const size_t cols = (size_t)T * Kp;
    DBuf da, db, de, dc, dtm, dj, du;
    CK(da.alloc(cols * 8));
    CK(db.alloc(cols * 8));
    CK(de.alloc(cols * 8));
    CK(dc.alloc((size_t)T * 8));
    CK(dtm.alloc((size_t)std::max(T - 1, 1) * 16));
    CK(dj.alloc(sizeof(JobDev)));
    CK(du.alloc(256));
    CK(cudaMemset(du.p, 0, 256));
    CK(cudaMemset(de.p, 0, cols * 8));
    CK(cudaMemcpy2D(de.p, (size_t)Kp * 8, eMatGrid_t, (size_t)K * 8, (size_t)K * 8, T, cudaMemcpyHostToDevice));
    if (T > 1) CK(cudaMemcpy(dtm.p, tm, (size_t)(T - 1) * 16, cudaMemcpyHostToDevice));
    JobDev J;
    std::memset(&J, 0, sizeof(J));
    J.alpha = (double*)da.p;
    J.beta = (double*)db.p;
    J.eG = (double*)de.p;
    J.c = (double*)dc.p;
    J.tm = (const double*)dtm.p;
    J.underflow = (int32_t*)du.p;
    CK(cudaMemcpy(dj.p, &J, sizeof(J), cudaMemcpyHostToDevice));
    BatchParams P;
    std::memset(&P, 0, sizeof(P));
    P.K = K;
    P.Kp = Kp;
    P.T = T;
    P.NH = 1;
    P.one_over_K = 1 / double(K);
    if (geo.CL == 2) {
        k_fb_generic<512, 16><<<dim3(1, 1), 512, 0, g_stream>>>(P, (const JobDev*)dj.p, 1);
        LAUNCHED_K("k_fb_generic");
        rc = QUILT_OK;
    } else {
        rc = with_geo(geo, [&](auto nt, auto ept) {
            k_fb_generic<decltype(nt)::value, decltype(ept)::value><<<dim3(1, 1), decltype(nt)::value, 0, g_stream>>>(P, (const JobDev*)dj.p, 1);
            LAUNCHED_K("k_fb_generic");
            return QUILT_OK;
        });
    }
    if (rc != QUILT_OK) return rc;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(g_stream));
    CK(cudaMemcpy2D(alphaHat_t, (size_t)K * 8, da.p, (size_t)Kp * 8, (size_t)K * 8, T, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy2D(betaHat_t, (size_t)K * 8, db.p, (size_t)Kp * 8, (size_t)K * 8, T, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(c, dc.p, (size_t)T * 8, cudaMemcpyDeviceToHost));
    return QUILT_OK;
}

}  // extern "C"
