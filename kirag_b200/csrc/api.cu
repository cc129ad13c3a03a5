// api.cu — the C ABI (include/kirag_b200.h) and the search orchestration.
//
// Boundary being replaced: the `faiss` module object used by
// /root/reference/retriever/index.py (IndexFlatIP ctor :13,23; add :32;
// search :47; write_index :62; read_index :73; ntotal :74,79) and the
// matmul+topk of /root/reference/knowledge_graph/models.py:1532-1538.
//
// Search pipeline per query chunk (KIRAG_PATH_AUTO):
//   1. convert queries to the bf16 swizzled block layout (+ ||q||)
//   2. geometric levels over the (permuted) 128-row shadow tiles:
//        tcgen05 filter scan  -> append (approx score, row) with score >= tau[q]
//        select               -> keep the k' best, tighten tau[q]
//   3. rescore the survivors that can still reach the top-k from the fp32 master (canonical order)
//   4. final sort by (score desc, id asc), write D/I, evaluate the certificate
//      (3 + 4 + the last select are ONE cluster kernel for calls of at most 128 queries)
//   5. queries whose buffer overflowed are re-run with a gentle level schedule, queries whose
//      certificate failed get one more bf16 pass with the provable threshold, whatever is left the
//      exact fp32 scan (resolve_chunk).
// Steps 1-4 never synchronise with the host (kirag_index_search_async, CUDA-graph capturable);
// step 5 follows the one synchronisation that reads the per-query flags.
#include "common.cuh"
#include "../../include/kirag_b200.h"

#include <cuda.h>  // driver-API TYPES only: the entry points are fetched with cudaGetDriverEntryPoint (no -lcuda)

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <mutex>
#include <new>
#include <vector>

namespace kirag {

static thread_local char g_err[768] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Chained (programmatic dependent) launches pay off where the kernels of a search are short: measured
// 12% at 430k rows, 4% at 2.6M rows x 32 queries, nothing at 1024 queries, and a slight LOSS at 4096
// queries (thousands of early-resident CTAs next to the tensor-bound scan).  So the attribute is only
// set for calls of at most kPdlMaxQueries queries.
constexpr int64_t kPdlMaxQueries = 1024;
static thread_local bool g_pdl_this_call = true;
bool pdl_enabled() {
    static const bool on = [] {
        const char* v = getenv("KIRAG_PDL");
        return !(v && *v && atoi(v) == 0);
    }();
    return on && g_pdl_this_call;
}

struct SmemAttrEntry { const void* kernel; int device; size_t bytes; };
static std::vector<SmemAttrEntry> g_smem_attr;
static std::mutex g_smem_attr_mu;
int ensure_dynamic_smem_impl(const void* kernel, size_t bytes) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    std::lock_guard<std::mutex> lock(g_smem_attr_mu);
    for (auto& e : g_smem_attr) {
        if (e.kernel == kernel && e.device == dev) {
            if (bytes <= e.bytes) return 0;
            cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
            if (err != cudaSuccess) {
                set_error("cudaFuncSetAttribute(MaxDynamicSharedMemorySize=%zu) failed: %s", bytes, cudaGetErrorString(err));
                return 1;
            }
            e.bytes = bytes;
            return 0;
        }
    }
    cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (err != cudaSuccess) {
        set_error("cudaFuncSetAttribute(MaxDynamicSharedMemorySize=%zu) failed: %s", bytes, cudaGetErrorString(err));
        return 1;
    }
    g_smem_attr.push_back({kernel, dev, bytes});
    return 0;
}

// Optional per-kernel timing of the dominant kernel (the tcgen05 filter scan): CUDA events
// recorded on the launching stream around every scan launch while enabled.  bench.py uses
// it for the roofline line; it is off by default.
struct ScanProfile {
    bool on = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;
    std::vector<double> launch_rows;
    std::vector<std::pair<int, cudaEvent_t>> marks;  // (tag, event) timeline of the search phases
    double rows = 0.0;  // corpus rows streamed by the recorded launches
    double queries = 0.0;
};
static ScanProfile g_prof;

// timeline tags: 0 start, 1 after convert/init, 10+l after scan of level l, 30+l after select of
// level l, 50 after rescore, 51 after final, 52 after certificate readback
static void prof_mark(int tag, cudaStream_t st) {
    if (!g_prof.on) return;
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, st);
    g_prof.marks.emplace_back(tag, e);
}

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    int ensure(size_t need) {
        if (need <= bytes) return 0;
        if (p) { cudaFree(p); p = nullptr; bytes = 0; }
        // round up so that slowly growing requests do not reallocate every call
        size_t want = (need + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            p = nullptr;
            set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
            return 1;
        }
        bytes = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; }
        ok = (cudaSetDevice(dev) == cudaSuccess);
        if (!ok) set_error("cudaSetDevice(%d) failed: %s", dev, cudaGetErrorString(cudaGetLastError()));
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

__global__ void fill_f32_kernel(float* p, float v, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
static int fill_f32(float* p, float v, int64_t n, cudaStream_t st) {
    if (n <= 0) return 0;
    fill_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, v, n);
    KIRAG_LAUNCH_OK("fill_f32_kernel");
    return 0;
}

// tau = -inf for real queries, +inf for the pad (pad queries never pass); counters and overflow flags zeroed
__global__ void init_search_state_kernel(float* tau, int* cnt, int* overflow, int64_t nq, int64_t nq_pad) {
    pdl_wait();  // the buffers may still be in use by the previous search's kernels
    pdl_launch_dependents();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq_pad) tau[i] = (i < nq) ? -INFINITY : INFINITY;
    if (i < nq) { cnt[i] = 0; overflow[i] = 0; }
}

__global__ void gather_rows_kernel(const float* __restrict__ src, const int* __restrict__ idx, int d,
                                   float* __restrict__ dst) {
    const int64_t r = blockIdx.x;
    const float* s = src + (int64_t)idx[r] * d;
    float* t = dst + r * d;
    for (int c = threadIdx.x; c < d; c += blockDim.x) t[c] = s[c];
}

}  // namespace kirag

using namespace kirag;

// ------------------------------------------------------------- growable device storage ---
// The corpus (fp32 master + bf16 shadow: 129 GB at 21M x 1024) must be able to GROW — the reference's build path
// calls index.add() once per 1M-row file with no reserve hook (faiss_index_corpus.py:42-46) — without ever holding
// two copies: cudaMalloc(new) + copy + cudaFree(old) needs old + new resident at once, which does not fit a 180 GB
// GPU beyond ~40 % fill.  So each buffer is a reserved VIRTUAL address range into which physical
// chunks are mapped as the index grows (cuMemAddressReserve / cuMemCreate / cuMemMap): growing maps one more
// chunk at the end, nothing is copied, the base address never changes.
struct VmmApi {
    bool ok = false;
    CUresult (*AddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*AddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*Create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
    CUresult (*Release)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*Map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*Unmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*SetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    CUresult (*GetGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
};
static const VmmApi& vmm_api() {
    static const VmmApi api = [] {
        VmmApi a;
        const char* off = getenv("KIRAG_NO_VMM");
        if (off && *off && atoi(off) != 0) return a;
        auto get = [](const char* name, void** fn) {
            cudaDriverEntryPointQueryResult qr;
            return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &qr) == cudaSuccess &&
                   qr == cudaDriverEntryPointSuccess && *fn != nullptr;
        };
        a.ok = get("cuMemAddressReserve", (void**)&a.AddressReserve) && get("cuMemAddressFree", (void**)&a.AddressFree) &&
               get("cuMemCreate", (void**)&a.Create) && get("cuMemRelease", (void**)&a.Release) &&
               get("cuMemMap", (void**)&a.Map) && get("cuMemUnmap", (void**)&a.Unmap) &&
               get("cuMemSetAccess", (void**)&a.SetAccess) &&
               get("cuMemGetAllocationGranularity", (void**)&a.GetGranularity);
        cudaGetLastError();
        return a;
    }();
    return api;
}

struct VmBuffer {
    CUdeviceptr base = 0;
    size_t reserved = 0, mapped = 0, gran = 0;
    int device = 0;
    struct Chunk { CUmemGenericAllocationHandle h; size_t off, size; };
    std::vector<Chunk> chunks;

    bool active() const { return base != 0; }
    void* ptr() const { return reinterpret_cast<void*>(base); }

    CUmemAllocationProp prop() const {
        CUmemAllocationProp p = {};
        p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        p.location.id = device;
        return p;
    }
    // reserve `bytes` of address space (no physical memory yet)
    int reserve(int dev, size_t bytes) {
        const VmmApi& a = vmm_api();
        if (!a.ok) return 1;
        device = dev;
        const CUmemAllocationProp p = prop();
        if (a.GetGranularity(&gran, &p, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || gran == 0) return 1;
        reserved = (bytes + gran - 1) / gran * gran;
        if (a.AddressReserve(&base, reserved, 0, 0, 0) != CUDA_SUCCESS) { base = 0; reserved = 0; return 1; }
        return 0;
    }
    // make [0, bytes) usable; 0 ok, 1 out of (physical or reserved) memory — nothing is changed then
    int ensure(size_t bytes) {
        if (bytes <= mapped) return 0;
        const VmmApi& a = vmm_api();
        const size_t want = (bytes + gran - 1) / gran * gran;
        if (want > reserved) return 1;
        Chunk c;
        c.off = mapped;
        c.size = want - mapped;
        const CUmemAllocationProp p = prop();
        if (a.Create(&c.h, c.size, &p, 0) != CUDA_SUCCESS) return 1;
        if (a.Map(base + c.off, c.size, 0, c.h, 0) != CUDA_SUCCESS) { a.Release(c.h); return 1; }
        CUmemAccessDesc acc = {};
        acc.location = p.location;
        acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        if (a.SetAccess(base + c.off, c.size, &acc, 1) != CUDA_SUCCESS) {
            a.Unmap(base + c.off, c.size);
            a.Release(c.h);
            return 1;
        }
        chunks.push_back(c);
        mapped = want;
        return 0;
    }
    void release() {
        const VmmApi& a = vmm_api();
        for (auto& c : chunks) { a.Unmap(base + c.off, c.size); a.Release(c.h); }
        chunks.clear();
        if (base) a.AddressFree(base, reserved);
        base = 0;
        reserved = mapped = 0;
    }
};

struct FastParams {
    int kprime;
    int growth_override;  // KIRAG_LEVEL_GROWTH (0: automatic)
    int cap_override;     // KIRAG_CAND_CAP (0: automatic)
    int first_growth_override;  // KIRAG_LEVEL1_GROWTH (0: automatic)
    int few_queries;            // the call has at most kFewQueries queries (set per call)
};

struct SearchCounters {
    int64_t n_fast = 0, n_exact = 0, n_cert_fail = 0, n_overflow = 0, n_rescan = 0, n_retry = 0, n_changed = 0;
    int levels = 0;
};

// An asynchronous search (kirag_index_search_async) whose certificate flags have not been examined yet.
struct PendingSearch {
    bool active = false;
    bool ev_recorded = false;  // false when the call was captured into a CUDA graph (events cannot be waited on then)
    bool fast_ok = false;
    int mode = 0, path = 0, k = 0;
    int64_t nq = 0, id_offset = 0;
    const float* qd = nullptr;
    float* Dd = nullptr;
    int64_t* Id = nullptr;
    cudaStream_t st = nullptr;
    long long launches0 = 0;
    SearchCounters counters;
    FastParams fp{};
};

struct kirag_index {
    int d = 0;
    int metric = 0;
    int device = 0;
    int num_sms = 148;
    int64_t ntotal = 0;
    int64_t capacity = 0;       // rows
    float* master = nullptr;    // [capacity, d] fp32
    uint8_t* shadow = nullptr;  // bf16 tiles, capacity rounded up to 128 rows (null if d % 64)
    VmBuffer vm_master, vm_shadow;  // where master / shadow live when virtual-memory growth is available
    bool use_vmm = false;
    unsigned* maxnorm2_bits = nullptr;  // device: [0] max ||y||^2, [1] max ||y - bf16(y)||^2, [2] max ||x||^2 (float bits)
    float maxnorm = 0.f;                // max_j ||y_j||, y = x - center (= x without a centre): what the shadow holds
    float maxerr = 0.f;                 // max_j ||y_j - bf16(y_j)||
    float maxnorm_x = 0.f;              // max_j ||x_j||
    float* center = nullptr;            // device [d]: the shadow stores bf16(x - center); null = not centred
    float center_norm = 0.f;
    bool center_decided = false;
    // workspaces (grow-only)
    DevBuf q_dev, D_dev, I_dev, qshadow, qnorm, cand, cnt, tau, tauk, overflow, flags, rescored;
    DevBuf dense, stage_a, stage_b, qmap, qsel, qnorm2, D_tmp, I_tmp, t_dev, center_sums;
    int* host_flags = nullptr;  // pinned: certificate read-back without a staging copy
    size_t host_flags_n = 0;
    PendingSearch pending;
    PendingSearch captured;  // the asynchronous search most recently captured into a CUDA graph (kirag_index_search_rearm)
    bool has_captured = false;
    cudaEvent_t pending_ev = nullptr;
};

static int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    return atoi(v);
}

// Centring decision (once per index, when it first holds kCenterMinRows rows): c = mean of the first rows, used only
// if it is a LARGE common component (||c||^2 >= 1/16 of the mean squared row norm).  i.i.d.-like corpora (||mean|| ~
// 1/sqrt(n)) stay uncentred and bit-identical to an index without this feature; E5-style embeddings (random-pair
// cosine 0.7+) get c ~ their common direction.  Any fixed c is valid — it shifts all scores of a query by <q, c> — so
// rows added later need not share the mean for exactness, only for the certificate to stay tight.
constexpr int64_t kCenterMinRows = 4096;
// E||mean||^2 of n i.i.d. unit rows is 1/n, so 8192 rows decide safely, and a centre that is 1 % off costs the
// certificate nothing; summing 65536 rows took 50 us of every transient kirag_topk_ip index.
constexpr int64_t kCenterSampleRows = 8192;
static int decide_center(kirag_index* h, int64_t rows_available, cudaStream_t st) {
    if (env_int("KIRAG_NO_CENTER", 0)) return 0;
    const int d = h->d;
    const int64_t rows = rows_available < kCenterSampleRows ? rows_available : kCenterSampleRows;
    DevBuf& sums = h->center_sums;  // kept: a cudaMalloc / cudaFree pair per transient index (kirag_topk_ip) cost ~60 us
    if (sums.ensure((size_t)(d + 1) * 4)) return 1;
    std::vector<float> host((size_t)d + 1);
    int rc = 1;
    do {
        if (cudaMemsetAsync(sums.p, 0, (size_t)(d + 1) * 4, st) != cudaSuccess) break;
        if (launch_column_sums(h->master, rows, d, sums.as<float>(), st)) break;
        if (cudaMemcpyAsync(host.data(), sums.p, (size_t)(d + 1) * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess) {
            set_error("decide_center: reading the column sums failed: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        double c2 = 0.0;
        for (int c = 0; c < d; ++c) {
            host[(size_t)c] = (float)((double)host[(size_t)c] / (double)rows);
            c2 += (double)host[(size_t)c] * (double)host[(size_t)c];
        }
        const double mean_sq = (double)host[(size_t)d] / (double)rows;
        rc = 0;
        if (!(c2 >= mean_sq / 16.0) || !(mean_sq > 0.0) || !(c2 == c2)) break;  // no large common component
        rc = 1;
        if (cudaMalloc((void**)&h->center, (size_t)d * 4) != cudaSuccess) {
            h->center = nullptr;
            set_error("decide_center: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        if (cudaMemcpyAsync(h->center, host.data(), (size_t)d * 4, cudaMemcpyHostToDevice, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess) {
            set_error("decide_center: uploading the centre failed");
            break;
        }
        h->center_norm = (float)sqrt(c2);
        rc = 0;
    } while (0);
    return rc;
}

static int index_grow(kirag_index* h, int64_t need_rows, cudaStream_t st) {
    if (need_rows <= h->capacity) return 0;
    const bool want_shadow = scan_tc_supported(h->d) != 0;
    const size_t row_m = (size_t)h->d * sizeof(float), row_s = (size_t)h->d * 2;
    if (h->capacity == 0 && !h->master && vmm_api().ok) {
        // address space for the largest index this GPU could ever hold (the whole device memory as fp32 rows)
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && total_b > 0) {
            const size_t max_rows = (size_t)round_up((int64_t)(total_b / row_m) + kTileRows, kTileRows);
            if (h->vm_master.reserve(h->device, max_rows * row_m) == 0 &&
                (!want_shadow || h->vm_shadow.reserve(h->device, max_rows * row_s) == 0)) {
                h->use_vmm = true;
                h->master = static_cast<float*>(h->vm_master.ptr());
                h->shadow = want_shadow ? static_cast<uint8_t*>(h->vm_shadow.ptr()) : nullptr;
            } else {
                h->vm_master.release();
                h->vm_shadow.release();
            }
        }
        cudaGetLastError();
    }
    // grow by a quarter beyond what is needed (amortises the mapping calls of many small adds); when that does
    // not fit any more, by exactly what is needed
    int64_t cap = round_up(need_rows > h->capacity + h->capacity / 4 ? need_rows : h->capacity + h->capacity / 4, kTileRows);
    const int64_t exact = round_up(need_rows, kTileRows);
    if (h->use_vmm) {
        for (int attempt = 0; attempt < 2; ++attempt) {
            const int64_t c = attempt == 0 ? cap : exact;
            // nothing has to be undone on failure: a mapped-but-unused tail of the master is just capacity
            if (h->vm_master.ensure((size_t)c * row_m) == 0 && (!want_shadow || h->vm_shadow.ensure((size_t)c * row_s) == 0)) {
                h->capacity = c;
                return 0;
            }
            if (cap == exact) break;
        }
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        set_error("index storage for %lld rows x %d (fp32 master%s) does not fit: %zu MiB free of %zu MiB",
                  (long long)exact, h->d, want_shadow ? " + bf16 shadow" : "", free_b >> 20, total_b >> 20);
        cudaGetLastError();
        return 1;
    }
    // no virtual-memory API (KIRAG_NO_VMM=1 / old driver): cudaMalloc + copy, one buffer after the other so that
    // at most ONE buffer exists twice at any time
    auto regrow = [&](void** buf, size_t row_bytes, size_t live_rows, const char* what) -> int {
        void* nb = nullptr;
        int64_t c = cap;
        cudaError_t e = cudaMalloc(&nb, (size_t)c * row_bytes);
        if (e != cudaSuccess && cap != exact) { cudaGetLastError(); c = exact; e = cudaMalloc(&nb, (size_t)c * row_bytes); }
        if (e != cudaSuccess) {
            set_error("cudaMalloc of the %s (%zu bytes for %lld rows) failed: %s", what, (size_t)c * row_bytes, (long long)c,
                      cudaGetErrorString(e));
            cudaGetLastError();
            return 1;
        }
        if (*buf && live_rows > 0) {
            if (cudaMemcpyAsync(nb, *buf, live_rows * row_bytes, cudaMemcpyDeviceToDevice, st) != cudaSuccess ||
                cudaStreamSynchronize(st) != cudaSuccess) {
                cudaFree(nb);
                set_error("device copy while growing the %s failed: %s", what, cudaGetErrorString(cudaGetLastError()));
                return 1;
            }
        }
        if (*buf) cudaFree(*buf);
        *buf = nb;
        cap = c;  // the second buffer must not be larger than the first
        return 0;
    };
    if (regrow((void**)&h->master, row_m, (size_t)h->ntotal, "fp32 master")) return 1;
    if (want_shadow && regrow((void**)&h->shadow, row_s, (size_t)round_up(h->ntotal, kTileRows), "bf16 shadow")) return 1;
    h->capacity = cap;
    return 0;
}

// ------------------------------------------------------------ exact path ----
// Answers queries qsub[0..nsub) (device, contiguous [nsub, d]) with the exact
// fp32 scan.  Output row of query i is out_row0 + i, or qmap_dev[i] if given.
static int exact_search(kirag_index* h, const float* qsub, int64_t nsub, int k, float* D, int64_t* I,
                        int64_t out_row0, const int* qmap_dev, int64_t id_offset, cudaStream_t st) {
    const int64_t n = h->ntotal;
    const int d = h->d;
    const int m = (int)((k < n) ? k : n);  // at most n real results
    const int64_t ld = round_up(n, 32);
    if (h->dense.ensure((size_t)kExactNQ * ld * sizeof(float))) return 1;
    const int64_t n_seg = (n + kSelectSeg - 1) / kSelectSeg;
    if (h->stage_a.ensure((size_t)kExactNQ * n_seg * m * sizeof(Cand))) return 1;
    for (int64_t g0 = 0; g0 < nsub; g0 += kExactNQ) {
        const int gq = (int)((nsub - g0 < kExactNQ) ? (nsub - g0) : kExactNQ);
        if (launch_scan_exact(h->master, n, d, qsub + g0 * d, gq, h->dense.as<float>(), ld, h->num_sms, st)) return 1;
        int nseg_i = 0;
        if (launch_select_dense(h->dense.as<float>(), ld, n, gq, m, h->stage_a.as<Cand>(), &nseg_i, st)) return 1;
        Cand* cur = h->stage_a.as<Cand>();
        int64_t cur_stride = (int64_t)nseg_i * m;
        int cur_count = (int)cur_stride;
        bool use_b = true;
        while (cur_count > kSelectSeg) {
            const int ns2 = (cur_count + kSelectSeg - 1) / kSelectSeg;
            DevBuf& dstb = use_b ? h->stage_b : h->stage_a;
            // stage_a is only reused after it has been fully consumed
            if (dstb.ensure((size_t)kExactNQ * ns2 * m * sizeof(Cand))) return 1;
            Cand* dst = dstb.as<Cand>();
            if (launch_select_pairs(cur, cur_stride, nullptr, cur_count, 0x7fffffff, gq, m, dst,
                                    (int64_t)ns2 * m, ns2, nullptr, nullptr, nullptr, st)) return 1;
            cur = dst;
            cur_stride = (int64_t)ns2 * m;
            cur_count = (int)cur_stride;
            use_b = !use_b;
        }
        if (launch_final(cur, cur_stride, nullptr, nullptr, cur_count, cur_count, gq, k,
                         qmap_dev ? D : D + (out_row0 + g0) * k, qmap_dev ? I : I + (out_row0 + g0) * k,
                         id_offset, nullptr, nullptr, nullptr, nullptr, 0.f, 0.f, 0, nullptr, nullptr,
                         qmap_dev ? qmap_dev + g0 : nullptr, st)) return 1;
    }
    return 0;
}

// ------------------------------------------------------------- fast path ----
// Level schedule of the filter path (tiles of 128 rows, walked in a pseudo-random order):
//   level 0   cap/2 rows, everything is appended (tau = -inf)
//   level 1   grows the prefix 4x: tau comes from few distinct tiles and is a noisy estimate when rows
//             are clustered by position, so the first step keeps 16x headroom in the buffer
//   then      L equal geometric steps with growth g <= g_max = cap / (4 k') (expected survivors of a
//             level = k' * g, 4x headroom), L the smallest count that reaches the end: equal steps
//             append the fewest survivors in total for a given number of launches
// Small query batches (<= kWideCapMaxQueries) get a 4x larger candidate buffer: the filter has slack
// there (HBM-bound), launches are what costs, and a 4x larger g_max removes two to three levels.
constexpr int64_t kWideCapMaxQueries = 128;
constexpr int64_t kQChunk = 16384;  // queries per pass of the search workspaces
constexpr int64_t kFewQueries = 8;
constexpr int kMaxGrowth = 32;


static bool fast_eligible(const kirag_index* h, int k, FastParams* fp) {
    if (!h->shadow || !scan_tc_supported(h->d)) return false;
    if (h->ntotal > 0x7fffff00LL) return false;
    int64_t kp = (int64_t)4 * k;  // over-fetch k' = 4k (north_star), at most 2048 (k <= 2048 is checked by the caller)
    if (kp < 32) kp = 32;
    if (kp > 2048) kp = 2048;
    if (kp < k) return false;
    fp->kprime = (int)kp;
    fp->growth_override = env_int("KIRAG_LEVEL_GROWTH", 0);
    fp->cap_override = env_int("KIRAG_CAND_CAP", 0);
    fp->first_growth_override = env_int("KIRAG_LEVEL1_GROWTH", 0);
    if (fp->cap_override > 0 && fp->cap_override < 4 * kp) return false;
    return true;
}

static int pick_cap(const FastParams& fp, int64_t nq) {
    int cap = (nq <= kWideCapMaxQueries) ? kWideCap : kSelectSeg;
    if (fp.cap_override > 0) cap = fp.cap_override;
    if (cap > kWideCap) cap = kWideCap;
    return cap;
}

// upper tile bound (exclusive) of every level
static std::vector<int64_t> level_bounds(int64_t n_tiles, int cap, const FastParams& fp) {
    std::vector<int64_t> hi;
    int64_t t = (cap / 2) / kTileRows;
    if (t < 1) t = 1;
    if (t > n_tiles) t = n_tiles;
    hi.push_back(t);
    if (t >= n_tiles) return hi;
    int gmax = cap / (4 * fp.kprime);
    if (gmax > kMaxGrowth) gmax = kMaxGrowth;
    if (fp.growth_override > 0) gmax = fp.growth_override;
    if (gmax < 2) gmax = 2;
    // growth of level 1: 4; 16 for calls of at most kFewQueries queries (KiRAG's own shape: 1-2 queries per
    // retrieval), where one level less is worth ~35 us per call on a 2.6M-row shard (0.92 -> 0.88 ms).  With 32 or
    // more queries the 4x more survivors of a 16x level cost more than the level saves (measured: 0.92 -> 0.95 ms
    // at 32 queries, profiles/r02_ab/r2f_knobs.log).  An overflow is re-answered with the gentle schedule.
    int g1 = (cap >= kWideCap && fp.few_queries) ? 16 : 4;
    if (fp.first_growth_override > 0) g1 = fp.first_growth_override;
    if (g1 > gmax) g1 = gmax;
    t = t * g1;
    if (t > n_tiles) t = n_tiles;
    hi.push_back(t);
    if (t >= n_tiles) return hi;
    const double ratio = (double)n_tiles / (double)t;
    int L = (int)ceil(log(ratio) / log((double)gmax) - 1e-9);
    if (L < 1) L = 1;
    const double g = pow(ratio, 1.0 / L);
    const int64_t base = t;
    for (int i = 1; i <= L; ++i) {
        int64_t b = (i == L) ? n_tiles : (int64_t)ceil((double)base * pow(g, (double)i));
        if (b <= hi.back()) b = hi.back() + 1;
        if (b > n_tiles) b = n_tiles;
        hi.push_back(b);
        if (b >= n_tiles) break;
    }
    return hi;
}

static int64_t pick_tile_mult(int64_t n_tiles) {
    // odd-ish multiplier near the golden ratio, coprime with n_tiles: consecutive
    // positions of the tile walk land far apart, so every level is a spread-out
    // sample of the corpus (robust against corpora sorted/clustered by position)
    if (n_tiles <= 2) return 1;
    int64_t m = (int64_t)(0.6180339887498949 * (double)n_tiles);
    if (m < 1) m = 1;
    auto gcd = [](int64_t a, int64_t b) { while (b) { int64_t t = a % b; a = b; b = t; } return a; };
    while (gcd(m, n_tiles) != 1) ++m;
    return m % n_tiles;
}

// Exactness certificate, the bound on |approx - canonical| for any corpus row j and query q.  With
// x~ = bf16(x), q~ = bf16(q):  <x~,q~> - <x,q> = <x~ - x, q~> + <x, q~ - q>, so
//   |.| <= ||x_j - x~_j|| (||q|| + ||q - q~||) + ||x_j|| ||q - q~||
// plus the fp32 accumulation error of both dot products (d * 2^-21 ||x|| ||q|| is generous: the bf16
// products are exact in fp32, sequential rounding costs at most d * 2^-24 per unit of ||x|| ||q||, a
// truncating accumulator twice that).  The rounding-error norms are MEASURED when rows / queries
// are converted (convert.cu), not bounded by 2^-9 ||x||: for typical data they are ~0.58 * 2^-9 ||x||,
// which makes eps ~1.4x tighter than the worst case.  eps(q) = a * ||q|| + b * ||q - q~||, 0.2 % slack
// for the fp32 evaluation of the norms themselves.
struct CertEps { float a, b; };
static CertEps cert_eps(const kirag_index* h) {
    // y = fl(x - c) is what the shadow rounds to bf16 (y = x, c = 0 without a centre):
    //   <q, x> = <q, y> + <q, c> + <q, (x - c) - y>                      |last term| <= 2^-24 Y ||q||
    //   |<q~, y~> - <q, y>| <= E_y (||q|| + ||q - q~||) + Y ||q - q~||     (bf16 rounding of both sides)
    //   fp32 accumulation: d 2^-22 Y ||q|| for the tensor-core dot, d 2^-22 X ||q|| for the canonical dot
    //   <q, c> is evaluated in fp32 by one warp: error <= 64 * 2^-24 ||q|| ||c||
    const double yn = (double)h->maxnorm, ey = (double)h->maxerr, xn = (double)h->maxnorm_x, cn = (double)h->center_norm;
    const double acc = (double)h->d * ldexp(1.0, -22);
    CertEps e;
    e.a = (float)((ey + acc * (yn + xn) + ldexp(1.0, -24) * yn + 64.0 * ldexp(1.0, -24) * cn) * 1.002);
    e.b = (float)((yn + 2.0 * ey) * 1.002);
    return e;
}

static int fast_search(kirag_index* h, const float* qd, int64_t nq, int k, float* D, int64_t* I,
                       int64_t id_offset, const FastParams& fp, int check_cert, int* levels_out,
                       cudaStream_t st) {
    const int d = h->d;
    const int64_t n = h->ntotal;
    struct PdlScope {
        explicit PdlScope(bool on) { g_pdl_this_call = on; }
        ~PdlScope() { g_pdl_this_call = true; }
    } pdl_scope(nq <= kPdlMaxQueries);
    ScanTcPlan plan;
    if (scan_tc_pick(nq, d, &plan)) return 1;
    const size_t qs_bytes = scan_tc_qshadow_bytes(nq, d, plan);
    if (h->qshadow.ensure(qs_bytes)) return 1;
    if (h->qnorm.ensure((size_t)nq * 12)) return 1;  // [nq] norms, [nq] bf16 rounding-error norms, [nq] <q, center>
    const int cap = pick_cap(fp, nq);
    if (h->cand.ensure((size_t)nq * cap * sizeof(Cand))) return 1;
    if (h->cnt.ensure((size_t)nq * 4)) return 1;
    const int64_t nq_pad = round_up(nq, 256);
    if (h->tau.ensure((size_t)nq_pad * 4)) return 1;
    if (h->overflow.ensure((size_t)nq * 4)) return 1;
    if (h->flags.ensure((size_t)nq * 4)) return 1;
    if (h->rescored.ensure((size_t)nq * fp.kprime * 4)) return 1;
    if (h->tauk.ensure((size_t)nq * 4)) return 1;
    // rescoring skips candidates more than 2 eps below the k-th best approximate score (KIRAG_RESCORE_ALL=1: all k')
    const bool skip_far = env_int("KIRAG_RESCORE_ALL", 0) == 0;
    // small batches (wide candidate buffer): last compaction + rescoring + final sort are ONE cluster kernel
    const bool fuse_tail = cap == kWideCap && fp.kprime <= 2048 && env_int("KIRAG_FUSED_TAIL", 1) != 0;
    if ((nq % plan.bq) != 0)  // only the pad rows of the last query tile need zeroing
        KIRAG_CUDA_OK(cudaMemsetAsync(h->qshadow.p, 0, qs_bytes, st));
    KIRAG_CUDA_OK(launch_chained(init_search_state_kernel, dim3((unsigned)((nq_pad + 255) / 256)), dim3(256), 0, st,
                                 h->tau.as<float>(), h->cnt.as<int>(), h->overflow.as<int>(), nq, nq_pad));
    KIRAG_LAUNCH_OK("init_search_state_kernel");
    prof_mark(0, st);
    float* const qcdot = h->qnorm.as<float>() + 2 * nq;
    if (launch_convert_rows(qd, nq, d, 0, h->qshadow.p, plan.q_tile_rows, nullptr, h->qnorm.as<float>(),
                            h->qnorm.as<float>() + nq, h->center, qcdot, st)) return 1;
    prof_mark(1, st);

    const int64_t n_tiles = (n + kTileRows - 1) / kTileRows;
    const int64_t mult = pick_tile_mult(n_tiles);
    FastParams fpl = fp;
    fpl.few_queries = nq <= kFewQueries ? 1 : 0;
    const std::vector<int64_t> bounds = level_bounds(n_tiles, cap, fpl);
    int64_t lo = 0;
    int levels = 0;
    for (const int64_t hi : bounds) {
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (g_prof.on) {
            KIRAG_CUDA_OK(cudaEventCreate(&e0));
            KIRAG_CUDA_OK(cudaEventCreate(&e1));
            KIRAG_CUDA_OK(cudaEventRecord(e0, st));
        }
        if (launch_scan_tc(h->shadow, n, d, h->qshadow.p, nq, plan, lo, hi, n_tiles, mult,
                           h->tau.as<float>(), h->cand.as<Cand>(), h->cnt.as<int>(), cap, h->num_sms,
                           levels == 0 ? 1 : 0, st)) return 1;
        if (g_prof.on) {
            KIRAG_CUDA_OK(cudaEventRecord(e1, st));
            g_prof.ev.emplace_back(e0, e1);
            int64_t r1 = hi * kTileRows; if (r1 > n) r1 = n;
            g_prof.rows += (double)(r1 - lo * kTileRows);
            g_prof.launch_rows.push_back((double)(r1 - lo * kTileRows));
            g_prof.queries = (double)nq;
        }
        prof_mark(10 + levels, st);
        const bool last = (hi == bounds.back());
        if (last && fuse_tail) break;  // the fused tail kernel below does this level's compaction itself
        if (launch_compact_topm(h->cand.as<Cand>(), cap, h->cnt.as<int>(), cap, (int)nq, fp.kprime,
                                h->tau.as<float>(), h->overflow.as<int>(), (last && skip_far) ? h->tauk.as<float>() : nullptr,
                                k, st)) return 1;
        prof_mark(30 + levels, st);
        ++levels;
        lo = hi;
    }
    const CertEps ce = cert_eps(h);
    if (fuse_tail) {
        ++levels;
        if (launch_tail_fused(h->cand.as<Cand>(), cap, h->cnt.as<int>(), cap, (int)nq, fp.kprime, h->tau.as<float>(),
                              h->overflow.as<int>(), h->master, d, qd, h->rescored.as<float>(), k, D, I, id_offset,
                              h->qnorm.as<float>(), h->qnorm.as<float>() + nq, qcdot, ce.a, ce.b, h->flags.as<int>(),
                              h->num_sms, st)) return 1;
        prof_mark(51, st);
        if (levels_out) *levels_out = levels;
        return 0;
    }
    if (launch_rescore(h->master, d, qd, h->cand.as<Cand>(), h->cnt.as<int>(), cap, fp.kprime,
                       h->rescored.as<float>(), nq, skip_far ? h->tauk.as<float>() : nullptr, h->qnorm.as<float>(),
                       h->qnorm.as<float>() + nq, ce.a, ce.b, st)) return 1;
    prof_mark(50, st);
    if (launch_final(h->cand.as<Cand>(), cap, h->rescored.as<float>(), h->cnt.as<int>(), 0, fp.kprime,
                     (int)nq, k, D, I, id_offset, h->tau.as<float>(), h->qnorm.as<float>(), h->qnorm.as<float>() + nq,
                     qcdot, ce.a, ce.b, check_cert, h->overflow.as<int>(), h->flags.as<int>(), nullptr, st)) return 1;
    prof_mark(51, st);
    if (levels_out) *levels_out = levels;
    return 0;
}

// tau2[i] = (k-th canonical score of flagged query i) - eps_i : every row of the true top-k has an
// approximate score >= tau2 (certificate); a query without k results gets -inf
__global__ void rescan_threshold_kernel(const float* __restrict__ D, const int* __restrict__ qmap,
                                        const float* __restrict__ qnorm, const float* __restrict__ qerr,
                                        const float* __restrict__ qcdot, int k,
                                        float eps_a, float eps_b, int64_t nb,
                                        int64_t nb_pad, float* __restrict__ tau2, int* __restrict__ cnt,
                                        int* __restrict__ overflow) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nb_pad) return;
    if (i >= nb) { tau2[i] = INFINITY; return; }
    const int q = qmap[i];
    const float kth = D[(int64_t)q * k + (k - 1)];
    // the shadow holds x - c: its scores are lower than the true ones by <q, c>
    tau2[i] = (kth > -FLT_MAX) ? kth - (eps_a * qnorm[q] + eps_b * qerr[q]) - qcdot[q] : -INFINITY;
    cnt[i] = 0;
    overflow[i] = 0;
}

// Second chance for queries whose certificate failed: ONE more pass of the bf16 filter over the whole
// shadow with the provable threshold tau2 = s_k - eps (fixed, no levels), then every collected
// candidate is rescored in fp32.  Exact by construction; only a buffer overflow (more than `cap`
// rows within eps of the k-th score) is left to the exact fp32 scan.  qsel: the flagged queries
// gathered contiguously; qmap: their row numbers in D/I.  still_bad (host) receives the overflowed ones.
static int rescan_search(kirag_index* h, const float* qsel, const float* qnorm_all, const float* qerr_all,
                         const float* qcdot_all, int64_t nb,
                         int k, float* D,
                         int64_t* I, const int* qmap_dev, int64_t id_offset, int cap, std::vector<int>* still_bad,
                         cudaStream_t st) {
    const int d = h->d;
    const int64_t n = h->ntotal;
    ScanTcPlan plan;
    if (scan_tc_pick(nb, d, &plan)) return 1;
    const size_t qs_bytes = scan_tc_qshadow_bytes(nb, d, plan);
    const int64_t nb_pad = round_up(nb, 256);
    if (h->qshadow.ensure(qs_bytes)) return 1;
    if (h->cand.ensure((size_t)nb * cap * sizeof(Cand))) return 1;
    if (h->cnt.ensure((size_t)nb_pad * 4) || h->tau.ensure((size_t)nb_pad * 4) || h->overflow.ensure((size_t)nb_pad * 4)) return 1;
    if (h->flags.ensure((size_t)nb_pad * 4)) return 1;
    if (h->rescored.ensure((size_t)nb * cap * 4)) return 1;
    const CertEps ce = cert_eps(h);
    // thresholds from the first-attempt results (still in D), before anything is overwritten
    rescan_threshold_kernel<<<(unsigned)((nb_pad + 255) / 256), 256, 0, st>>>(
        D, qmap_dev, qnorm_all, qerr_all, qcdot_all, k, ce.a, ce.b, nb, nb_pad, h->tau.as<float>(), h->cnt.as<int>(),
        h->overflow.as<int>());
    KIRAG_LAUNCH_OK("rescan_threshold_kernel");
    if ((nb % plan.bq) != 0) KIRAG_CUDA_OK(cudaMemsetAsync(h->qshadow.p, 0, qs_bytes, st));
    if (launch_convert_rows(qsel, nb, d, 0, h->qshadow.p, plan.q_tile_rows, nullptr, nullptr, nullptr, nullptr, nullptr,
                            st)) return 1;
    const int64_t n_tiles = (n + kTileRows - 1) / kTileRows;
    if (launch_scan_tc(h->shadow, n, d, h->qshadow.p, nb, plan, 0, n_tiles, n_tiles, 1, h->tau.as<float>(),
                       h->cand.as<Cand>(), h->cnt.as<int>(), cap, h->num_sms, 1, st)) return 1;
    if (launch_rescore(h->master, d, qsel, h->cand.as<Cand>(), h->cnt.as<int>(), cap, cap, h->rescored.as<float>(), nb,
                       nullptr, nullptr, nullptr, 0.f, 0.f, st)) return 1;
    // overflow = appended count beyond the buffer
    std::vector<int> counts((size_t)nb);
    KIRAG_CUDA_OK(cudaMemcpyAsync(counts.data(), h->cnt.p, (size_t)nb * 4, cudaMemcpyDeviceToHost, st));
    if (launch_final(h->cand.as<Cand>(), cap, h->rescored.as<float>(), h->cnt.as<int>(), 0, cap, (int)nb, k, D, I,
                     id_offset, nullptr, nullptr, nullptr, nullptr, 0.f, 0.f, 0, nullptr, nullptr, qmap_dev, st)) return 1;
    KIRAG_CUDA_OK(cudaStreamSynchronize(st));
    for (int64_t i = 0; i < nb; ++i)
        if (counts[(size_t)i] > cap) still_bad->push_back((int)i);
    return 0;
}

// ------------------------------------------------ enqueue / resolve of one chunk ----
// A search is split in two halves so that the certificate never forces a host synchronisation in the
// middle of the device work:
//   enqueue_chunk  every kernel of the first attempt (filter levels, rescoring, final sort + certificate) and
//                  the copy of the per-query certificate flags into pinned host memory.  No synchronisation,
//                  no pageable copies: stream-ordered and CUDA-graph capturable (workspaces must be warm).
//   resolve_chunk  AFTER the stream has been synchronised by whoever needs the results: looks at the flags and
//                  re-answers the flagged queries (second bf16 pass with the provable threshold, then the exact
//                  fp32 scan for whatever is left).  Flagged queries are rare (none on i.i.d. data).

static int ensure_host_flags(kirag_index* h, int64_t cq) {
    if (h->host_flags_n >= (size_t)cq) return 0;
    if (h->host_flags) cudaFreeHost(h->host_flags);
    h->host_flags = nullptr;
    h->host_flags_n = 0;
    const size_t want = (size_t)round_up(cq, 4096);
    KIRAG_CUDA_OK(cudaHostAlloc((void**)&h->host_flags, want * sizeof(int), cudaHostAllocDefault));
    h->host_flags_n = want;
    return 0;
}

// mode: 0 = nothing to verify (empty index / exact path), 1 = filter path, flags are on their way to host_flags
static int enqueue_chunk(kirag_index* h, const float* qd, int64_t cq, int k, float* Dd, int64_t* Id, int64_t id_offset,
                         bool fast_ok, const FastParams& fp, SearchCounters* c, int* mode, cudaStream_t st) {
    *mode = 0;
    if (h->ntotal == 0) return launch_fill_pad(Dd, Id, cq * k, st);
    if (!fast_ok) {
        if (exact_search(h, qd, cq, k, Dd, Id, 0, nullptr, id_offset, st)) return 1;
        c->n_exact += cq;
        return 0;
    }
    if (ensure_host_flags(h, cq)) return 1;  // may synchronise the device: before anything is enqueued
    if (fast_search(h, qd, cq, k, Dd, Id, id_offset, fp, 1, &c->levels, st)) return 1;
    // one read-back: flags[i] = 0 ok, 1 certificate failed, 2 candidate buffer overflowed
    KIRAG_CUDA_OK(cudaMemcpyAsync(h->host_flags, h->flags.p, (size_t)cq * 4, cudaMemcpyDeviceToHost, st));
    *mode = 1;
    return 0;
}

__global__ void scatter_rows_kernel(const float* __restrict__ Ds, const int64_t* __restrict__ Is, const int* __restrict__ qmap,
                                    int k, float* __restrict__ D, int64_t* __restrict__ I) {
    const int64_t src = blockIdx.x, dst = qmap[src];
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
        D[dst * k + j] = Ds[src * k + j];
        I[dst * k + j] = Is[src * k + j];
    }
}

// The stream must have been synchronised since enqueue_chunk.  Returns with the stream idle.
// Escalation ladder for the flagged queries of a chunk (KIRAG_PATH_AUTO):
//   overflowed candidate buffer  -> the filter path once more with the gentle level schedule (growth 2, then 4):
//                                   the first attempt's schedule is tuned for latency and bets on the sampled
//                                   thresholds being representative; corpora clustered by position can lose that bet
//   certificate failed           -> ONE more bf16 pass with the provable threshold s_k - eps
//   whatever is left             -> the exact fp32 scan
static int resolve_chunk(kirag_index* h, const float* qd, int64_t cq, int k, float* Dd, int64_t* Id, int64_t id_offset,
                         int path, const FastParams& fp, SearchCounters* c, cudaStream_t st) {
    const int d = h->d;
    const int* flags = h->host_flags;
    std::vector<int> cert, ovf;
    for (int64_t i = 0; i < cq; ++i) {
        if (flags[(size_t)i] == 2) { ovf.push_back((int)i); ++c->n_overflow; }
        else if (flags[(size_t)i]) { cert.push_back((int)i); ++c->n_cert_fail; }
    }
    const int64_t n_bad = (int64_t)(cert.size() + ovf.size());
    if (n_bad == 0 || path != KIRAG_PATH_AUTO) {
        c->n_fast += cq;
        return 0;
    }
    // the norm / flag buffers are reused below: keep the first attempt's norms (+ error norms, + <q, center>)
    if (h->qnorm2.ensure((size_t)cq * 12)) return 1;
    KIRAG_CUDA_OK(cudaMemcpyAsync(h->qnorm2.p, h->qnorm.p, (size_t)cq * 12, cudaMemcpyDeviceToDevice, st));
    std::vector<int> exact;
    const bool no_rescan = env_int("KIRAG_NO_RESCAN", 0) != 0;
    const bool gentle_already = fp.growth_override > 0 && fp.growth_override <= 4 && fp.first_growth_override > 0;
    if (!ovf.empty() && !gentle_already && !env_int("KIRAG_NO_RETRY", 0)) {
        const int64_t nb = (int64_t)ovf.size();
        if (h->qmap.ensure((size_t)nb * 4) || h->qsel.ensure((size_t)nb * d * 4)) return 1;
        if (h->D_tmp.ensure((size_t)nb * k * 4) || h->I_tmp.ensure((size_t)nb * k * 8)) return 1;
        KIRAG_CUDA_OK(cudaMemcpyAsync(h->qmap.p, ovf.data(), (size_t)nb * 4, cudaMemcpyHostToDevice, st));
        gather_rows_kernel<<<(unsigned)nb, 256, 0, st>>>(qd, h->qmap.as<int>(), d, h->qsel.as<float>());
        KIRAG_LAUNCH_OK("gather_rows_kernel");
        FastParams gentle = fp;
        gentle.growth_override = 4;
        gentle.first_growth_override = 2;
        if (fast_search(h, h->qsel.as<float>(), nb, k, h->D_tmp.as<float>(), h->I_tmp.as<int64_t>(), id_offset, gentle, 1,
                        nullptr, st)) return 1;
        std::vector<int> f2((size_t)nb);
        KIRAG_CUDA_OK(cudaMemcpyAsync(f2.data(), h->flags.p, (size_t)nb * 4, cudaMemcpyDeviceToHost, st));
        scatter_rows_kernel<<<(unsigned)nb, 128, 0, st>>>(h->D_tmp.as<float>(), h->I_tmp.as<int64_t>(), h->qmap.as<int>(), k,
                                                          Dd, Id);
        KIRAG_LAUNCH_OK("scatter_rows_kernel");
        KIRAG_CUDA_OK(cudaStreamSynchronize(st));
        for (int64_t i = 0; i < nb; ++i) {
            if (f2[(size_t)i] == 0) ++c->n_retry;
            else if (f2[(size_t)i] == 2) exact.push_back(ovf[(size_t)i]);
            else cert.push_back(ovf[(size_t)i]);  // its rows of D now hold a valid k-th score for the rescan threshold
        }
    } else {
        exact = ovf;
    }
    if (!cert.empty() && no_rescan) {
        exact.insert(exact.end(), cert.begin(), cert.end());
        cert.clear();
    }
    if (!cert.empty()) {
        const int64_t nb = (int64_t)cert.size();
        if (h->qmap.ensure((size_t)nb * 4)) return 1;
        if (h->qsel.ensure((size_t)nb * d * 4)) return 1;
        KIRAG_CUDA_OK(cudaMemcpyAsync(h->qmap.p, cert.data(), (size_t)nb * 4, cudaMemcpyHostToDevice, st));
        gather_rows_kernel<<<(unsigned)nb, 256, 0, st>>>(qd, h->qmap.as<int>(), d, h->qsel.as<float>());
        KIRAG_LAUNCH_OK("gather_rows_kernel");
        std::vector<int> still_bad;
        if (rescan_search(h, h->qsel.as<float>(), h->qnorm2.as<float>(), h->qnorm2.as<float>() + cq,
                          h->qnorm2.as<float>() + 2 * cq, nb, k, Dd, Id, h->qmap.as<int>(), id_offset, kSelectSeg, &still_bad,
                          st)) return 1;
        for (int i : still_bad) exact.push_back(cert[(size_t)i]);
        c->n_rescan += nb - (int64_t)still_bad.size();
    }
    if (!exact.empty()) {
        const int64_t nb = (int64_t)exact.size();
        if (h->qmap.ensure((size_t)nb * 4)) return 1;
        if (h->qsel.ensure((size_t)nb * d * 4)) return 1;
        KIRAG_CUDA_OK(cudaMemcpyAsync(h->qmap.p, exact.data(), (size_t)nb * 4, cudaMemcpyHostToDevice, st));
        gather_rows_kernel<<<(unsigned)nb, 256, 0, st>>>(qd, h->qmap.as<int>(), d, h->qsel.as<float>());
        KIRAG_LAUNCH_OK("gather_rows_kernel");
        if (exact_search(h, h->qsel.as<float>(), nb, k, Dd, Id, 0, h->qmap.as<int>(), id_offset, st)) return 1;
        c->n_exact += nb;
    }
    // the index vectors live on this stack frame until the copies above have been consumed
    KIRAG_CUDA_OK(cudaStreamSynchronize(st));
    c->n_fast += cq - n_bad;
    c->n_changed += n_bad;
    return 0;
}

static void fill_stats(kirag_search_stats_t* stats, int64_t nq, const SearchCounters& c, bool fast_ok, int path,
                       long long launches0) {
    if (!stats) return;
    stats->nq = nq;
    stats->n_fast = c.n_fast;
    stats->n_exact = c.n_exact;
    stats->n_cert_fail = c.n_cert_fail;
    stats->n_overflow = c.n_overflow;
    stats->n_rescan = c.n_rescan;
    stats->n_retry = c.n_retry;
    stats->levels = c.levels;
    stats->path = fast_ok ? path : KIRAG_PATH_EXACT;
    stats->kernel_launches = g_launches.load() - launches0;
}

// Completes an asynchronous search whose certificate has not been looked at yet (kirag_index_search_async).
static int finish_pending(kirag_index* h, kirag_search_stats_t* stats, int64_t* n_changed) {
    if (n_changed) *n_changed = 0;
    PendingSearch& p = h->pending;
    if (!p.active) {
        if (stats) memset(stats, 0, sizeof(*stats));
        return 0;
    }
    p.active = false;
    DeviceGuard guard(h->device);
    if (!guard.ok) return 1;
    if (p.ev_recorded) KIRAG_CUDA_OK(cudaEventSynchronize(h->pending_ev));
    else KIRAG_CUDA_OK(cudaStreamSynchronize(p.st));
    SearchCounters c = p.counters;
    if (p.mode == 1 && resolve_chunk(h, p.qd, p.nq, p.k, p.Dd, p.Id, p.id_offset, p.path, p.fp, &c, p.st)) return 1;
    fill_stats(stats, p.nq, c, p.fast_ok, p.path, p.launches0);
    if (n_changed) *n_changed = c.n_changed;
    return 0;
}

static int search_impl(kirag_index* h, const float* q, int64_t nq, int k, float* D, int64_t* I,
                       int ptrs_are_device, int64_t id_offset, int path, kirag_search_stats_t* stats,
                       cudaStream_t st) {
    KIRAG_CHECK(h != nullptr, "search: null index");
    KIRAG_CHECK(k > 0, "search: k must be positive (got %d)", k);
    KIRAG_CHECK(k <= 2048, "search: k=%d exceeds the supported maximum of 2048", k);
    KIRAG_CHECK(nq >= 0, "search: negative nq");
    KIRAG_CHECK(path >= 0 && path <= 2, "search: unknown path %d", path);
    if (finish_pending(h, nullptr, nullptr)) return 1;  // an unfinished asynchronous search owns the workspaces
    if (stats) memset(stats, 0, sizeof(*stats));
    if (nq == 0) return 0;
    KIRAG_CHECK(q && D && I, "search: null buffer");
    DeviceGuard guard(h->device);
    if (!guard.ok) return 1;
    const long long launches0 = g_launches.load();
    const int d = h->d;
    SearchCounters c;
    FastParams fp{};
    const bool fast_ok = (path != KIRAG_PATH_EXACT) && h->ntotal > 0 && fast_eligible(h, k, &fp);

    for (int64_t q0 = 0; q0 < nq; q0 += kQChunk) {
        const int64_t cq = (nq - q0 < kQChunk) ? (nq - q0) : kQChunk;
        const float* qd;
        float* Dd;
        int64_t* Id;
        if (ptrs_are_device) {
            qd = q + q0 * d; Dd = D + q0 * k; Id = I + q0 * k;
        } else {
            if (h->q_dev.ensure((size_t)cq * d * 4)) return 1;
            if (h->D_dev.ensure((size_t)cq * k * 4)) return 1;
            if (h->I_dev.ensure((size_t)cq * k * 8)) return 1;
            KIRAG_CUDA_OK(cudaMemcpyAsync(h->q_dev.p, q + q0 * d, (size_t)cq * d * 4, cudaMemcpyHostToDevice, st));
            qd = h->q_dev.as<float>(); Dd = h->D_dev.as<float>(); Id = h->I_dev.as<int64_t>();
        }
        int mode = 0;
        if (enqueue_chunk(h, qd, cq, k, Dd, Id, id_offset, fast_ok, fp, &c, &mode, st)) return 1;
        // host buffers: the results travel with the certificate flags, ONE synchronisation per chunk
        if (!ptrs_are_device) {
            KIRAG_CUDA_OK(cudaMemcpyAsync(D + q0 * k, Dd, (size_t)cq * k * 4, cudaMemcpyDeviceToHost, st));
            KIRAG_CUDA_OK(cudaMemcpyAsync(I + q0 * k, Id, (size_t)cq * k * 8, cudaMemcpyDeviceToHost, st));
        }
        if (mode == 1 || !ptrs_are_device) KIRAG_CUDA_OK(cudaStreamSynchronize(st));
        if (mode == 1) {
            const int64_t changed0 = c.n_changed;
            if (resolve_chunk(h, qd, cq, k, Dd, Id, id_offset, path, fp, &c, st)) return 1;
            if (!ptrs_are_device && c.n_changed != changed0) {  // some rows were re-answered: fetch them again
                KIRAG_CUDA_OK(cudaMemcpyAsync(D + q0 * k, Dd, (size_t)cq * k * 4, cudaMemcpyDeviceToHost, st));
                KIRAG_CUDA_OK(cudaMemcpyAsync(I + q0 * k, Id, (size_t)cq * k * 8, cudaMemcpyDeviceToHost, st));
                KIRAG_CUDA_OK(cudaStreamSynchronize(st));
            }
        }
    }
    fill_stats(stats, nq, c, fast_ok, path, launches0);
    return 0;
}

// Rows [r0, r0 + n) of the master are in place: (re)decide the centre if it is time, convert them to the shadow
// and refresh the norm maxima.  One pass, one synchronisation.
static int finalize_rows(kirag_index* h, int64_t r0, int64_t n, cudaStream_t st) {
    if (h->shadow && !h->center_decided && r0 + n >= kCenterMinRows) {
        if (decide_center(h, r0 + n, st)) return 1;
        h->center_decided = true;
        if (h->center && r0 > 0) {
            KIRAG_CUDA_OK(cudaMemsetAsync(h->maxnorm2_bits, 0, 8, st));
            if (launch_convert_rows(h->master, r0, h->d, 0, h->shadow, kTileRows, h->maxnorm2_bits, nullptr, nullptr, h->center,
                                    nullptr, st)) return 1;
        }
    }
    if (launch_convert_rows(h->master + r0 * (int64_t)h->d, n, h->d, r0, h->shadow, kTileRows, h->maxnorm2_bits, nullptr,
                            nullptr, h->center, nullptr, st)) return 1;
    if (h->shadow && ((r0 + n) % kTileRows) != 0) {
        // rows of the last tile beyond ntotal: masked by the scan, but newly mapped memory is not zeroed — keep them
        // finite.  In the block layout rows rr..127 of a tile are the tail of each of its d/64 16-KB blocks.
        const int64_t end = r0 + n;
        const int64_t tile = end / kTileRows;
        const int rr = (int)(end - tile * kTileRows);
        uint8_t* tile_base = h->shadow + (size_t)tile * ((size_t)h->d * kTileRows * 2);
        KIRAG_CUDA_OK(cudaMemset2DAsync(tile_base + (size_t)rr * 128, (size_t)kTileRows * 128, 0, (size_t)(kTileRows - rr) * 128,
                                        (size_t)(h->d / kKChunk), st));
    }
    unsigned bits[3] = {0, 0, 0};
    KIRAG_CUDA_OK(cudaMemcpyAsync(bits, h->maxnorm2_bits, 12, cudaMemcpyDeviceToHost, st));
    KIRAG_CUDA_OK(cudaStreamSynchronize(st));
    float m2[3];
    memcpy(m2, bits, 12);
    h->maxnorm = sqrtf(m2[0]);
    h->maxerr = h->shadow ? sqrtf(m2[1]) : 0.f;
    h->maxnorm_x = h->shadow ? sqrtf(m2[2]) : h->maxnorm;
    h->ntotal = r0 + n;
    return 0;
}

// ================================================================= C ABI ====
extern "C" {

int kirag_abi_version(void) { return KIRAG_ABI_VERSION; }

const char* kirag_last_error(void) { return g_err; }

int kirag_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int kirag_profile_enable(int on) {
    for (auto& pr : g_prof.ev) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    g_prof.ev.clear();
    g_prof.launch_rows.clear();
    for (auto& mk : g_prof.marks) cudaEventDestroy(mk.second);
    g_prof.marks.clear();
    g_prof.rows = 0.0;
    g_prof.on = on != 0;
    return 0;
}

int kirag_profile_read_timeline(int* tags_out, double* ms_since_first, int64_t max_n) {
    // timeline of the recorded search phases (does not clear); ms are relative to the first mark
    int64_t i = 0;
    for (auto& mk : g_prof.marks) {
        if (i >= max_n) break;
        KIRAG_CUDA_OK(cudaEventSynchronize(mk.second));
        float t = 0.f;
        KIRAG_CUDA_OK(cudaEventElapsedTime(&t, g_prof.marks[0].second, mk.second));
        if (tags_out) tags_out[i] = mk.first;
        if (ms_since_first) ms_since_first[i] = t;
        ++i;
    }
    return (int)i;
}

int kirag_profile_read_launches(double* ms_out, double* rows_out, int64_t max_n) {
    // per-launch durations (does not clear; call before kirag_profile_read)
    int64_t i = 0;
    for (auto& pr : g_prof.ev) {
        if (i >= max_n) break;
        KIRAG_CUDA_OK(cudaEventSynchronize(pr.second));
        float t = 0.f;
        KIRAG_CUDA_OK(cudaEventElapsedTime(&t, pr.first, pr.second));
        if (ms_out) ms_out[i] = t;
        if (rows_out) rows_out[i] = g_prof.launch_rows[(size_t)i];
        ++i;
    }
    return 0;
}

int kirag_profile_read(double* scan_ms, int64_t* scan_launches, double* scan_rows) {
    double ms = 0.0;
    for (auto& pr : g_prof.ev) {
        KIRAG_CUDA_OK(cudaEventSynchronize(pr.second));
        float t = 0.f;
        KIRAG_CUDA_OK(cudaEventElapsedTime(&t, pr.first, pr.second));
        ms += t;
    }
    if (scan_ms) *scan_ms = ms;
    if (scan_launches) *scan_launches = (int64_t)g_prof.ev.size();
    if (scan_rows) *scan_rows = g_prof.rows;
    for (auto& pr : g_prof.ev) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    g_prof.ev.clear();
    g_prof.launch_rows.clear();
    for (auto& mk : g_prof.marks) cudaEventDestroy(mk.second);
    g_prof.marks.clear();
    g_prof.rows = 0.0;
    return 0;
}

int kirag_index_create(int d, int metric, int device, kirag_index_t** out) {
    KIRAG_CHECK(out != nullptr, "index_create: null out pointer");
    *out = nullptr;
    KIRAG_CHECK(d > 0 && d <= 16384, "index_create: unsupported dimension %d", d);
    KIRAG_CHECK(metric == KIRAG_METRIC_INNER_PRODUCT,
                "index_create: only the inner-product metric is implemented (got %d)", metric);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error("index_create: no CUDA device available (%s); this library has no CPU path",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        cudaGetLastError();
        return 1;
    }
    KIRAG_CHECK(device >= 0 && device < ndev, "index_create: device %d out of range [0,%d)", device, ndev);
    cudaDeviceProp prop;
    KIRAG_CUDA_OK(cudaGetDeviceProperties(&prop, device));
    KIRAG_CHECK(prop.major == 10, "index_create: device %d is sm_%d%d; this library is built for sm_100a only",
                device, prop.major, prop.minor);
    DeviceGuard guard(device);
    if (!guard.ok) return 1;
    kirag_index* h = new (std::nothrow) kirag_index();
    KIRAG_CHECK(h != nullptr, "index_create: out of host memory");
    h->d = d;
    h->metric = metric;
    h->device = device;
    h->num_sms = prop.multiProcessorCount;
    if (cudaMalloc((void**)&h->maxnorm2_bits, 12) != cudaSuccess || cudaMemset(h->maxnorm2_bits, 0, 12) != cudaSuccess) {
        set_error("index_create: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
        delete h;
        return 1;
    }
    *out = h;
    return 0;
}

int kirag_index_destroy(kirag_index_t* h) {
    if (!h) return 0;
    DeviceGuard guard(h->device);
    cudaDeviceSynchronize();
    if (h->use_vmm) {
        h->vm_master.release();
        h->vm_shadow.release();
    } else {
        if (h->master) cudaFree(h->master);
        if (h->shadow) cudaFree(h->shadow);
    }
    if (h->maxnorm2_bits) cudaFree(h->maxnorm2_bits);
    if (h->center) cudaFree(h->center);
    DevBuf* bufs[] = {&h->q_dev, &h->D_dev, &h->I_dev, &h->qshadow, &h->qnorm, &h->cand, &h->cnt, &h->tau, &h->tauk,
                      &h->overflow, &h->flags, &h->rescored, &h->dense, &h->stage_a, &h->stage_b, &h->qmap, &h->qsel, &h->qnorm2, &h->t_dev, &h->center_sums,
                      &h->D_tmp, &h->I_tmp};
    for (DevBuf* b : bufs) b->release();
    if (h->host_flags) cudaFreeHost(h->host_flags);
    if (h->pending_ev) cudaEventDestroy(h->pending_ev);
    delete h;
    return 0;
}

int kirag_index_reserve(kirag_index_t* h, int64_t n_total) {
    KIRAG_CHECK(h != nullptr, "index_reserve: null index");
    KIRAG_CHECK(n_total >= 0, "index_reserve: negative size");
    DeviceGuard guard(h->device);
    if (!guard.ok) return 1;
    return index_grow(h, n_total, 0);
}

int kirag_index_add(kirag_index_t* h, const float* x, int64_t n, int x_is_device, void* stream) {
    KIRAG_CHECK(h != nullptr, "index_add: null index");
    KIRAG_CHECK(n >= 0, "index_add: negative n");
    if (n == 0) return 0;
    KIRAG_CHECK(x != nullptr, "index_add: null data");
    DeviceGuard guard(h->device);
    if (!guard.ok) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    if (finish_pending(h, nullptr, nullptr)) return 1;
    KIRAG_CHECK(h->ntotal + n <= 0x7fffff00LL, "index_add: more than 2^31 rows per device are not supported");
    if (index_grow(h, h->ntotal + n, st)) return 1;
    float* dst = h->master + h->ntotal * (int64_t)h->d;
    KIRAG_CUDA_OK(cudaMemcpyAsync(dst, x, (size_t)n * h->d * sizeof(float),
                                  x_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    return finalize_rows(h, h->ntotal, n, st);
}

int64_t kirag_index_ntotal(const kirag_index_t* h) { return h ? h->ntotal : -1; }
int kirag_index_dim(const kirag_index_t* h) { return h ? h->d : -1; }

int kirag_index_search(kirag_index_t* h, const float* q, int64_t nq, int k, float* D, int64_t* I,
                       int ptrs_are_device, int64_t id_offset, void* stream) {
    int path = env_int("KIRAG_PATH", KIRAG_PATH_AUTO);
    return search_impl(h, q, nq, k, D, I, ptrs_are_device, id_offset, path, nullptr, (cudaStream_t)stream);
}

int kirag_index_search_ex(kirag_index_t* h, const float* q, int64_t nq, int k, float* D, int64_t* I,
                          int ptrs_are_device, int64_t id_offset, int path, kirag_search_stats_t* stats,
                          void* stream) {
    return search_impl(h, q, nq, k, D, I, ptrs_are_device, id_offset, path, stats, (cudaStream_t)stream);
}

int kirag_index_search_async(kirag_index_t* h, const float* q, int64_t nq, int k, float* D, int64_t* I,
                             int64_t id_offset, void* stream) {
    KIRAG_CHECK(h != nullptr, "search_async: null index");
    KIRAG_CHECK(k > 0 && k <= 2048, "search_async: k=%d not in [1, 2048]", k);
    KIRAG_CHECK(nq >= 0 && nq <= kQChunk, "search_async: nq=%lld not in [0, %lld] (split larger batches)", (long long)nq,
                (long long)kQChunk);
    if (finish_pending(h, nullptr, nullptr)) return 1;
    if (nq == 0) return 0;
    KIRAG_CHECK(q && D && I, "search_async: null buffer");
    DeviceGuard guard(h->device);
    if (!guard.ok) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    PendingSearch& p = h->pending;
    p = PendingSearch();
    p.launches0 = g_launches.load();
    p.fast_ok = h->ntotal > 0 && fast_eligible(h, k, &p.fp);
    p.path = KIRAG_PATH_AUTO;
    if (enqueue_chunk(h, q, nq, k, D, I, id_offset, p.fast_ok, p.fp, &p.counters, &p.mode, st)) return 1;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cap);
    if (cap == cudaStreamCaptureStatusNone) {
        if (!h->pending_ev) KIRAG_CUDA_OK(cudaEventCreateWithFlags(&h->pending_ev, cudaEventDisableTiming));
        KIRAG_CUDA_OK(cudaEventRecord(h->pending_ev, st));
        p.ev_recorded = true;
    }
    p.nq = nq; p.k = k; p.id_offset = id_offset;
    p.qd = q; p.Dd = D; p.Id = I; p.st = st;
    p.active = true;
    if (cap != cudaStreamCaptureStatusNone) {
        h->captured = p;
        h->has_captured = true;
    }
    return 0;
}

int kirag_index_search_rearm(kirag_index_t* h, void* stream) {
    KIRAG_CHECK(h != nullptr, "search_rearm: null index");
    KIRAG_CHECK(h->has_captured, "search_rearm: no asynchronous search has been captured on this index");
    if (finish_pending(h, nullptr, nullptr)) return 1;
    h->pending = h->captured;
    h->pending.st = (cudaStream_t)stream;
    h->pending.ev_recorded = false;
    h->pending.launches0 = g_launches.load();
    h->pending.active = true;
    return 0;
}

// FNV-1a over everything a captured search has baked into its kernel arguments
int kirag_index_state_token(const kirag_index_t* h, uint64_t* token) {
    KIRAG_CHECK(h != nullptr && token != nullptr, "state_token: null argument");
    uint64_t x = 1469598103934665603ULL;
    auto mix = [&x](uint64_t v) {
        for (int i = 0; i < 8; ++i) { x ^= (v >> (8 * i)) & 0xffu; x *= 1099511628211ULL; }
    };
    mix((uint64_t)h->ntotal);
    mix((uint64_t)(uintptr_t)h->master);
    mix((uint64_t)(uintptr_t)h->shadow);
    mix((uint64_t)(uintptr_t)h->center);
    uint32_t f[4];
    memcpy(&f[0], &h->maxnorm, 4); memcpy(&f[1], &h->maxerr, 4); memcpy(&f[2], &h->maxnorm_x, 4); memcpy(&f[3], &h->center_norm, 4);
    mix(((uint64_t)f[0] << 32) | f[1]);
    mix(((uint64_t)f[2] << 32) | f[3]);
    const DevBuf* bufs[] = {&h->qshadow, &h->qnorm, &h->cand, &h->cnt, &h->tau, &h->tauk, &h->overflow, &h->flags, &h->rescored};
    for (const DevBuf* b : bufs) { mix((uint64_t)(uintptr_t)b->p); mix((uint64_t)b->bytes); }
    mix((uint64_t)(uintptr_t)h->host_flags);
    *token = x;
    return 0;
}

int kirag_index_search_finish(kirag_index_t* h, kirag_search_stats_t* stats, int64_t* n_changed) {
    KIRAG_CHECK(h != nullptr, "search_finish: null index");
    return finish_pending(h, stats, n_changed);
}

int kirag_index_search_flags(const kirag_index_t* h, const int** flags_dev) {
    KIRAG_CHECK(h != nullptr && flags_dev != nullptr, "search_flags: null argument");
    *flags_dev = (h->pending.active && h->pending.mode == 1) ? h->flags.as<int>() : nullptr;
    return 0;
}

int kirag_index_reconstruct(const kirag_index_t* h, int64_t i0, int64_t n, float* out, int out_is_device,
                            void* stream) {
    KIRAG_CHECK(h != nullptr && out != nullptr, "index_reconstruct: null argument");
    KIRAG_CHECK(i0 >= 0 && n >= 0 && i0 + n <= h->ntotal, "index_reconstruct: rows [%lld,%lld) out of range (ntotal=%lld)",
                (long long)i0, (long long)(i0 + n), (long long)h->ntotal);
    if (n == 0) return 0;
    DeviceGuard guard(h->device);
    if (!guard.ok) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    KIRAG_CUDA_OK(cudaMemcpyAsync(out, h->master + i0 * (int64_t)h->d, (size_t)n * h->d * 4,
                                  out_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    if (!out_is_device) KIRAG_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

/* debug / test hook: dense approximate (bf16 tcgen05) scores, out is host [ntotal, nq] */
int kirag_index_debug_scores(kirag_index_t* h, const float* q_host, int64_t nq, float* out_host) {
    KIRAG_CHECK(h && q_host && out_host, "debug_scores: null argument");
    KIRAG_CHECK(h->shadow && scan_tc_supported(h->d), "debug_scores: index has no bf16 shadow (d=%d)", h->d);
    KIRAG_CHECK(nq > 0 && h->ntotal > 0, "debug_scores: empty input");
    DeviceGuard guard(h->device);
    if (!guard.ok) return 1;
    cudaStream_t st = 0;
    const int d = h->d;
    ScanTcPlan plan;
    if (scan_tc_pick(nq, d, &plan)) return 1;
    const char* force = getenv("KIRAG_DEBUG_BQ");
    if (force && *force) {
        // 32 / 64 / 128 / 256: single-CTA kernels (32 resident, the others streamed); 512: the 2-CTA kernel
        // with 256-query tiles (2512 / 4512: two / four pairs per cluster with multicast query blocks); 1128: the
        // 2-CTA kernel with a resident 128-query tile; 1064: resident 64 (single CTA); 2064: resident 64 (2-CTA)
        plan.bq = atoi(force);
        plan.pair = 0;
        plan.multi = 1;
        plan.resident = (plan.bq == 32) ? 1 : 0;
        if (plan.bq == 512) { plan.bq = 256; plan.pair = 1; }
        if (plan.bq == 2512) { plan.bq = 256; plan.pair = 1; plan.multi = 2; }  // two pairs per cluster, multicast queries
        if (plan.bq == 4512) { plan.bq = 256; plan.pair = 1; plan.multi = 4; }
        if (plan.bq == 1128) { plan.bq = 128; plan.pair = 1; plan.resident = 1; }
        if (plan.bq == 1064) { plan.bq = 64; plan.resident = 1; }
        if (plan.bq == 2064) { plan.bq = 64; plan.pair = 1; plan.resident = 1; }  // 2-CTA kernel, resident 64-query tile
        plan.q_tile_rows = plan.pair ? plan.bq / 2 : plan.bq;
    }
    const size_t qs_bytes = scan_tc_qshadow_bytes(nq, d, plan);
    const int64_t nq_pad = round_up(nq, 256);
    DevBuf qd, qs, tau, cnt, dump, cdot;
    std::vector<float> cdot_host;
    int rc = 1;
    do {
        if (cdot.ensure((size_t)nq_pad * 4)) break;
        if (qd.ensure((size_t)nq * d * 4) || qs.ensure(qs_bytes) || tau.ensure((size_t)nq_pad * 4) ||
            cnt.ensure((size_t)nq_pad * 4) || dump.ensure((size_t)h->ntotal * nq * 4)) break;
        if (cudaMemcpyAsync(qd.p, q_host, (size_t)nq * d * 4, cudaMemcpyHostToDevice, st) != cudaSuccess) break;
        if (cudaMemsetAsync(qs.p, 0, qs_bytes, st) != cudaSuccess) break;
        if (cudaMemsetAsync(cnt.p, 0, (size_t)nq_pad * 4, st) != cudaSuccess) break;
        if (cudaMemsetAsync(dump.p, 0, (size_t)h->ntotal * nq * 4, st) != cudaSuccess) break;
        if (fill_f32(tau.as<float>(), INFINITY, nq_pad, st)) break;
        if (launch_convert_rows(qd.as<float>(), nq, d, 0, qs.p, plan.q_tile_rows, nullptr, nullptr, nullptr, h->center,
                                h->center ? cdot.as<float>() : nullptr, st)) break;
        if (launch_scan_tc_dump(h->shadow, h->ntotal, d, qs.p, nq, plan, tau.as<float>(), cnt.as<int>(),
                                dump.as<float>(), nq, h->num_sms, st)) break;
        if (h->center) {
            cdot_host.resize((size_t)nq);
            if (cudaMemcpyAsync(cdot_host.data(), cdot.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess) break;
        }
        if (cudaMemcpyAsync(out_host, dump.p, (size_t)h->ntotal * nq * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess) {
            set_error("debug_scores: copy/sync failed: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        // a centred shadow scores <q, x - c>: add <q, c> back so that the hook reports approximations of <q, x>
        if (h->center)
            for (int64_t r = 0; r < h->ntotal; ++r)
                for (int64_t qi = 0; qi < nq; ++qi) out_host[r * nq + qi] += cdot_host[(size_t)qi];
        rc = 0;
    } while (0);
    qd.release(); qs.release(); tau.release(); cnt.release(); dump.release(); cdot.release();
    return rc;
}

/* host-only: the level schedule fast_search() would use (no device needed) */
int kirag_debug_level_schedule(int64_t n_rows, int64_t nq, int k, int d, int64_t* rows_hi_out, int max_levels,
                               int* cap_out, int* kprime_out) {
    if (!(n_rows > 0 && nq > 0 && k > 0 && rows_hi_out && max_levels > 0)) {
        set_error("debug_level_schedule: bad argument");
        return -2;
    }
    int64_t kp = (int64_t)4 * k;
    if (kp < 32) kp = 32;
    if (kp > 2048) kp = 2048;
    if (kp < k || !scan_tc_supported(d) || n_rows > 0x7fffff00LL) return -1;  // not eligible for the filter path
    FastParams fp{};
    fp.kprime = (int)kp;
    fp.growth_override = env_int("KIRAG_LEVEL_GROWTH", 0);
    fp.cap_override = env_int("KIRAG_CAND_CAP", 0);
    fp.first_growth_override = env_int("KIRAG_LEVEL1_GROWTH", 0);
    if (fp.cap_override > 0 && fp.cap_override < 4 * kp) return -1;
    const int cap = pick_cap(fp, nq);
    fp.few_queries = nq <= kFewQueries ? 1 : 0;
    const int64_t n_tiles = (n_rows + kTileRows - 1) / kTileRows;
    const std::vector<int64_t> b = level_bounds(n_tiles, cap, fp);
    if ((int)b.size() > max_levels) {
        set_error("debug_level_schedule: %zu levels do not fit in %d slots", b.size(), max_levels);
        return -2;
    }
    for (size_t i = 0; i < b.size(); ++i) {
        const int64_t r = b[i] * kTileRows;
        rows_hi_out[i] = r < n_rows ? r : n_rows;
    }
    if (cap_out) *cap_out = cap;
    if (kprime_out) *kprime_out = (int)kp;
    return (int)b.size();
}

int kirag_index_device_ptrs(const kirag_index_t* h, const float** master_f32, const void** shadow_bf16) {
    KIRAG_CHECK(h != nullptr, "index_device_ptrs: null index");
    if (master_f32) *master_f32 = h->master;
    if (shadow_bf16) *shadow_bf16 = h->shadow;
    return 0;
}

// ---- FAISS "IxFI" container (faiss 1.8.0 index_write.cpp: write_index_header
//      + WRITEXBVECTOR(codes)); restated from the published format -------------
static const size_t kIoChunkRows = 65536;

int kirag_index_save(const kirag_index_t* h, const char* path) {
    KIRAG_CHECK(h != nullptr && path != nullptr, "index_save: null argument");
    DeviceGuard guard(h->device);
    if (!guard.ok) return 1;
    FILE* f = fopen(path, "wb");
    KIRAG_CHECK(f != nullptr, "index_save: cannot open %s for writing", path);
    bool ok = true;
    const uint32_t fourcc = (uint32_t)'I' | ((uint32_t)'x' << 8) | ((uint32_t)'F' << 16) | ((uint32_t)'I' << 24);
    const int32_t d = h->d;
    const int64_t ntotal = h->ntotal;
    const int64_t dummy = (int64_t)1 << 20;
    const uint8_t trained = 1;
    const int32_t metric = 0;  // METRIC_INNER_PRODUCT
    const uint64_t words = (uint64_t)ntotal * (uint64_t)d;
    ok = ok && fwrite(&fourcc, 4, 1, f) == 1 && fwrite(&d, 4, 1, f) == 1 && fwrite(&ntotal, 8, 1, f) == 1 &&
         fwrite(&dummy, 8, 1, f) == 1 && fwrite(&dummy, 8, 1, f) == 1 && fwrite(&trained, 1, 1, f) == 1 &&
         fwrite(&metric, 4, 1, f) == 1 && fwrite(&words, 8, 1, f) == 1;
    // two pinned buffers: the device -> host copy of piece i+1 runs while piece i is being written
    float* stage[2] = {nullptr, nullptr};
    cudaEvent_t copied[2] = {nullptr, nullptr};
    cudaStream_t st = nullptr;
    const size_t chunk_bytes = kIoChunkRows * (size_t)d * 4;
    if (ok && ntotal > 0) {
        bool alloc_ok = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess;
        for (int i = 0; i < 2; ++i)
            alloc_ok = alloc_ok && cudaMallocHost((void**)&stage[i], chunk_bytes) == cudaSuccess &&
                       cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming) == cudaSuccess;
        if (!alloc_ok) {
            set_error("index_save: pinned staging buffers: %s", cudaGetErrorString(cudaGetLastError()));
            ok = false;
        }
    }
    const int64_t n_pieces = (ntotal + (int64_t)kIoChunkRows - 1) / (int64_t)kIoChunkRows;
    auto piece_rows = [&](int64_t i) {
        const int64_t r = i * (int64_t)kIoChunkRows;
        return (ntotal - r < (int64_t)kIoChunkRows) ? (ntotal - r) : (int64_t)kIoChunkRows;
    };
    auto enqueue = [&](int64_t i) {
        const int b = (int)(i & 1);
        return cudaMemcpyAsync(stage[b], h->master + i * (int64_t)kIoChunkRows * d, (size_t)piece_rows(i) * d * 4,
                               cudaMemcpyDeviceToHost, st) == cudaSuccess &&
               cudaEventRecord(copied[b], st) == cudaSuccess;
    };
    if (ok && n_pieces > 0) ok = enqueue(0);
    for (int64_t i = 0; ok && i < n_pieces; ++i) {
        const int b = (int)(i & 1);
        if (cudaEventSynchronize(copied[b]) != cudaSuccess) { ok = false; break; }
        if (i + 1 < n_pieces && !enqueue(i + 1)) { ok = false; break; }  // the other buffer was written out last round
        const size_t n_el = (size_t)piece_rows(i) * d;
        ok = fwrite(stage[b], 4, n_el, f) == n_el;
    }
    if (st) cudaStreamSynchronize(st);
    for (int i = 0; i < 2; ++i) {
        if (stage[i]) cudaFreeHost(stage[i]);
        if (copied[i]) cudaEventDestroy(copied[i]);
    }
    if (st) cudaStreamDestroy(st);
    if (fclose(f) != 0) ok = false;
    KIRAG_CHECK(ok, "index_save: write to %s failed", path);
    return 0;
}

// faiss.read_index(path, IO_FLAG_MMAP) (retriever/index.py:73).  The corpus lives in HBM, so the file is streamed
// there once: storage for exactly ntotal rows is mapped up front, the payload is read in 64 MB pieces into two
// pinned buffers and copied while the next piece is being read (no per-piece synchronisation: the host only waits
// for the piece that used the same buffer two pieces ago), then ONE convert pass builds the bf16 shadow.  The
// IO_FLAG_MMAP the reference passes asks FAISS not to read the file eagerly; it has no meaning for a device-resident
// index and is accepted and ignored.
int kirag_index_load(const char* path, int device, kirag_index_t** out) {
    KIRAG_CHECK(path != nullptr && out != nullptr, "index_load: null argument");
    *out = nullptr;
    FILE* f = fopen(path, "rb");
    KIRAG_CHECK(f != nullptr, "index_load: cannot open %s", path);
    uint32_t fourcc = 0; int32_t d = 0; int64_t ntotal = 0, dummy = 0; uint8_t trained = 0; int32_t metric = 0;
    uint64_t words = 0;
    bool ok = fread(&fourcc, 4, 1, f) == 1 && fread(&d, 4, 1, f) == 1 && fread(&ntotal, 8, 1, f) == 1 &&
              fread(&dummy, 8, 1, f) == 1 && fread(&dummy, 8, 1, f) == 1 && fread(&trained, 1, 1, f) == 1 &&
              fread(&metric, 4, 1, f) == 1;
    const uint32_t ixfi = (uint32_t)'I' | ((uint32_t)'x' << 8) | ((uint32_t)'F' << 16) | ((uint32_t)'I' << 24);
    if (!ok || fourcc != ixfi) {
        fclose(f);
        set_error("index_load: %s is not an IxFI (IndexFlatIP) file (fourcc 0x%08x)", path, fourcc);
        return 1;
    }
    if (metric > 1) { float arg; ok = ok && fread(&arg, 4, 1, f) == 1; }
    ok = ok && fread(&words, 8, 1, f) == 1;
    if (!ok || metric != 0 || d <= 0 || ntotal < 0 || words != (uint64_t)ntotal * (uint64_t)d) {
        fclose(f);
        set_error("index_load: %s has an inconsistent header (d=%d ntotal=%lld metric=%d words=%llu)", path, d,
                  (long long)ntotal, metric, (unsigned long long)words);
        return 1;
    }
    kirag_index_t* h = nullptr;
    if (kirag_index_create(d, KIRAG_METRIC_INNER_PRODUCT, device, &h)) { fclose(f); return 1; }
    if (ntotal == 0) { fclose(f); *out = h; return 0; }
    DeviceGuard guard(device);
    float* stage[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    cudaStream_t st = nullptr;
    int rc = 1;
    do {
        if (!guard.ok) break;
        if (kirag_index_reserve(h, ntotal)) break;
        const size_t row_bytes = (size_t)d * 4;
        int64_t piece_rows = (int64_t)(((size_t)64 << 20) / row_bytes);
        if (piece_rows < 1) piece_rows = 1;
        if (piece_rows > ntotal) piece_rows = ntotal;
        if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) { set_error("index_load: cudaStreamCreate failed"); break; }
        bool alloc_ok = true;
        for (int i = 0; i < 2; ++i)
            alloc_ok = alloc_ok && cudaMallocHost((void**)&stage[i], (size_t)piece_rows * row_bytes) == cudaSuccess &&
                       cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming) == cudaSuccess;
        if (!alloc_ok) { set_error("index_load: pinned staging buffers: %s", cudaGetErrorString(cudaGetLastError())); break; }
        bool io_ok = true;
        int64_t piece = 0;
        for (int64_t r = 0; r < ntotal && io_ok; r += piece_rows, ++piece) {
            const int b = (int)(piece & 1);
            const int64_t rows = (ntotal - r < piece_rows) ? (ntotal - r) : piece_rows;
            if (piece >= 2 && cudaEventSynchronize(done[b]) != cudaSuccess) { io_ok = false; break; }  // buffer free again?
            if (fread(stage[b], row_bytes, (size_t)rows, f) != (size_t)rows) {
                set_error("index_load: %s is truncated", path);
                io_ok = false;
                break;
            }
            if (cudaMemcpyAsync(h->master + r * (int64_t)d, stage[b], (size_t)rows * row_bytes, cudaMemcpyHostToDevice, st) != cudaSuccess ||
                cudaEventRecord(done[b], st) != cudaSuccess) {
                set_error("index_load: host-to-device copy failed: %s", cudaGetErrorString(cudaGetLastError()));
                io_ok = false;
            }
        }
        if (!io_ok) { cudaStreamSynchronize(st); break; }
        if (finalize_rows(h, 0, ntotal, st)) break;
        rc = 0;
    } while (0);
    for (int i = 0; i < 2; ++i) {
        if (stage[i]) cudaFreeHost(stage[i]);
        if (done[i]) cudaEventDestroy(done[i]);
    }
    if (st) cudaStreamDestroy(st);
    fclose(f);
    if (rc) { kirag_index_destroy(h); return 1; }
    *out = h;
    return 0;
}

int kirag_merge_topk(const float* D_all, const int64_t* I_all, int G, int64_t nq, int k, float* D_out,
                     int64_t* I_out, int ptrs_are_device, int device, void* stream) {
    KIRAG_CHECK(G > 0 && k > 0 && nq >= 0, "merge_topk: bad shape G=%d nq=%lld k=%d", G, (long long)nq, k);
    if (nq == 0) return 0;
    KIRAG_CHECK(D_all && I_all && D_out && I_out, "merge_topk: null buffer");
    DeviceGuard guard(device);
    if (!guard.ok) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    if (ptrs_are_device) return launch_merge(D_all, I_all, G, nq, k, D_out, I_out, st);
    const size_t nin = (size_t)G * nq * k, nout = (size_t)nq * k;
    float* dD = nullptr; int64_t* dI = nullptr; float* oD = nullptr; int64_t* oI = nullptr;
    int rc = 1;
    do {
        if (cudaMalloc((void**)&dD, nin * 4) != cudaSuccess || cudaMalloc((void**)&dI, nin * 8) != cudaSuccess ||
            cudaMalloc((void**)&oD, nout * 4) != cudaSuccess || cudaMalloc((void**)&oI, nout * 8) != cudaSuccess) {
            set_error("merge_topk: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        if (cudaMemcpyAsync(dD, D_all, nin * 4, cudaMemcpyHostToDevice, st) != cudaSuccess ||
            cudaMemcpyAsync(dI, I_all, nin * 8, cudaMemcpyHostToDevice, st) != cudaSuccess) {
            set_error("merge_topk: H2D copy failed");
            break;
        }
        if (launch_merge(dD, dI, G, nq, k, oD, oI, st)) break;
        if (cudaMemcpyAsync(D_out, oD, nout * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaMemcpyAsync(I_out, oI, nout * 8, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess) {
            set_error("merge_topk: D2H copy failed: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        rc = 0;
    } while (0);
    if (dD) cudaFree(dD); if (dI) cudaFree(dI); if (oD) cudaFree(oD); if (oI) cudaFree(oI);
    return rc;
}

// The transient candidate matrix of a call is indexed in a per-(device, d) scratch index that is kept
// between calls: its device buffers (fp32 copy, bf16 shadow, search workspaces) are allocated once and
// grow-only, so a call costs one copy + one convert pass over T and one search — no cudaMalloc / cudaFree /
// device-wide synchronisation per call.
struct ScratchIndex { int device; int d; kirag_index_t* h; };
static std::vector<ScratchIndex> g_scratch;
static std::mutex g_scratch_mu;

int kirag_topk_ip(const float* q, int64_t nq, const float* t, int64_t nt, int d, int k, float* D, int64_t* I,
                  int ptrs_are_device, int device, void* stream) {
    KIRAG_CHECK(nt >= 0 && nq >= 0, "topk_ip: negative size");
    std::lock_guard<std::mutex> lock(g_scratch_mu);  // calls on one (device, d) share the scratch index
    kirag_index_t* h = nullptr;
    for (auto& e : g_scratch)
        if (e.device == device && e.d == d) h = e.h;
    if (!h) {
        if (kirag_index_create(d, KIRAG_METRIC_INNER_PRODUCT, device, &h)) return 1;
        g_scratch.push_back({device, d, h});
    }
    DeviceGuard guard(device);
    if (!guard.ok) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    if (finish_pending(h, nullptr, nullptr)) return 1;
    // KiRAG's own shape (models.py:1514-1542: 1-2 chain queries against the 10^2-10^3 triples of the accumulated
    // documents): building a transient index (copy, bf16 shadow, norm read-back = two host synchronisations) costs far
    // more than the contraction.  The fp32 scan + select run directly on the caller's matrix instead: three
    // stream-ordered launches per four queries, the same canonical scores and (score desc, id asc) order.
    // (measured, tools/aligner_latency.py: 2 x 1000 78 us here; above one select segment of 8192 rows the transient
    // index wins — 8 x 20000: 439 us in place, 16 x 20000 through the index 233 us)
    if (nq > 0 && nq <= 2 * kExactNQ && nt > 0 && nt <= kSelectSeg) {
        KIRAG_CHECK(q && t && D && I, "topk_ip: null buffer");
        KIRAG_CHECK(k > 0 && k <= 2048, "topk_ip: k=%d not in [1, 2048]", k);
        const float* qd = q;
        const float* td = t;
        float* Dd = D;
        int64_t* Id = I;
        if (!ptrs_are_device) {
            if (h->t_dev.ensure((size_t)nt * d * 4) || h->q_dev.ensure((size_t)nq * d * 4) ||
                h->D_dev.ensure((size_t)nq * k * 4) || h->I_dev.ensure((size_t)nq * k * 8)) return 1;
            KIRAG_CUDA_OK(cudaMemcpyAsync(h->t_dev.p, t, (size_t)nt * d * 4, cudaMemcpyHostToDevice, st));
            KIRAG_CUDA_OK(cudaMemcpyAsync(h->q_dev.p, q, (size_t)nq * d * 4, cudaMemcpyHostToDevice, st));
            qd = h->q_dev.as<float>(); td = h->t_dev.as<float>(); Dd = h->D_dev.as<float>(); Id = h->I_dev.as<int64_t>();
        }
        float* const own_master = h->master;
        const int64_t own_n = h->ntotal;
        h->master = const_cast<float*>(td);  // borrowed for the launches below only
        h->ntotal = nt;
        const int rc_small = exact_search(h, qd, nq, k, Dd, Id, 0, nullptr, 0, st);
        h->master = own_master;
        h->ntotal = own_n;
        if (rc_small) return 1;
        if (!ptrs_are_device) {
            KIRAG_CUDA_OK(cudaMemcpyAsync(D, Dd, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, st));
            KIRAG_CUDA_OK(cudaMemcpyAsync(I, Id, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, st));
            KIRAG_CUDA_OK(cudaStreamSynchronize(st));
        }
        return 0;
    }
    // forget the previous call's rows (capacity and workspaces stay)
    h->ntotal = 0;
    h->maxnorm = h->maxerr = h->maxnorm_x = h->center_norm = 0.f;
    h->center_decided = false;
    if (h->center) { cudaFree(h->center); h->center = nullptr; }
    KIRAG_CUDA_OK(cudaMemsetAsync(h->maxnorm2_bits, 0, 12, st));
    int rc = 0;
    if (nt > 0 && ptrs_are_device && (reinterpret_cast<uintptr_t>(t) & 15) == 0) {
        // device-resident candidates: the caller's matrix IS the fp32 master for the duration of this (synchronous)
        // call — no 4*nt*d-byte copy; only the bf16 shadow and the norm maxima are built in the scratch handle
        if (index_grow(h, nt, st)) return 1;
        float* const own_master = h->master;
        h->master = const_cast<float*>(t);
        rc = finalize_rows(h, 0, nt, st);
        if (!rc) rc = kirag_index_search(h, q, nq, k, D, I, ptrs_are_device, 0, stream);
        h->master = own_master;
        return rc;
    }
    if (nt > 0) rc = kirag_index_add(h, t, nt, ptrs_are_device, stream);
    if (!rc) rc = kirag_index_search(h, q, nq, k, D, I, ptrs_are_device, 0, stream);
    return rc;
}

int kirag_topk_ip_release(void) {
    std::lock_guard<std::mutex> lock(g_scratch_mu);
    for (auto& e : g_scratch) kirag_index_destroy(e.h);
    g_scratch.clear();
    return 0;
}

int kirag_pool_normalize(const void* hidden, const void* mask, float* out, int64_t B, int64_t S, int64_t H,
                         int64_t sb, int64_t ss, int64_t mb, int hidden_dtype, int mask_dtype, int mode,
                         int normalize, int device, void* stream) {
    KIRAG_CHECK(hidden && out, "pool_normalize: null buffer");
    DeviceGuard guard(device);
    if (!guard.ok) return 1;
    return launch_pool_normalize(hidden, mask, out, nullptr, nullptr, B, S, H, sb, ss, mb, hidden_dtype, mask_dtype, mode,
                                 normalize, (cudaStream_t)stream);
}

int kirag_pool_normalize_fwd_saved(const void* hidden, const void* mask, float* out, float* pooled_norm,
                                   int64_t B, int64_t S, int64_t H, int64_t sb, int64_t ss, int64_t mb,
                                   int hidden_dtype, int mask_dtype, int mode, int normalize, int device,
                                   void* stream) {
    KIRAG_CHECK(hidden && out && pooled_norm, "pool_normalize_fwd_saved: null buffer");
    DeviceGuard guard(device);
    if (!guard.ok) return 1;
    return launch_pool_normalize(hidden, mask, out, nullptr, pooled_norm, B, S, H, sb, ss, mb, hidden_dtype, mask_dtype,
                                 mode, normalize, (cudaStream_t)stream);
}

int kirag_pool_normalize_typed(const void* hidden, const void* mask, float* out, void* out_typed, float* pooled_norm,
                               int64_t B, int64_t S, int64_t H, int64_t sb, int64_t ss, int64_t mb,
                               int hidden_dtype, int mask_dtype, int mode, int normalize, int device,
                               void* stream) {
    KIRAG_CHECK(hidden && out && out_typed, "pool_normalize_typed: null buffer");
    DeviceGuard guard(device);
    if (!guard.ok) return 1;
    return launch_pool_normalize(hidden, mask, out, out_typed, pooled_norm, B, S, H, sb, ss, mb, hidden_dtype, mask_dtype,
                                 mode, normalize, (cudaStream_t)stream);
}

int kirag_pool_normalize_backward(const float* grad_out, const float* out, const float* pooled_norm,
                                  const void* mask, void* grad_hidden, int64_t B, int64_t S, int64_t H,
                                  int64_t mb, int hidden_dtype, int mask_dtype, int mode, int normalize,
                                  int device, void* stream) {
    KIRAG_CHECK(grad_out && out && pooled_norm && grad_hidden, "pool_normalize_backward: null buffer");
    DeviceGuard guard(device);
    if (!guard.ok) return 1;
    return launch_pool_normalize_backward(grad_out, out, pooled_norm, mask, grad_hidden, B, S, H, mb, hidden_dtype,
                                          mask_dtype, mode, normalize, (cudaStream_t)stream);
}

}  // extern "C"
