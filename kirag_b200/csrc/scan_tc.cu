// scan_tc.cu — the bf16 tcgen05 filter scan: the hot loop of the search.
//
// What it replaces: the inner loop of faiss.IndexFlatIP.search (reference call
// site /root/reference/retriever/index.py:47) — blocked sgemm of a query block
// against corpus blocks followed by a per-query heap/reservoir update.  Here
// the contraction runs on the 5th-generation tensor cores and the
// "does this score enter the running top-k'" test is fused into the
// accumulator read-out, so the score matrix never exists in memory.
//
// Data flow per CTA (persistent, one CTA per SM, 192 threads):
//   warp 0   producer : cp.async.bulk (TMA engine, SASS UBLKCP) of 16 KB shadow
//                       blocks [128 corpus rows x 64 k] into an N-stage ring;
//                       blocks are stored in HBM as the exact SWIZZLE_128B
//                       K-major shared-memory image, so each copy is one
//                       contiguous 16 KB read
//   warp 1   MMA      : tcgen05.mma.cta_group::1.kind::f16, M=128 (corpus rows)
//                       x N=BQ (queries) x K=16, bf16 in, fp32 accumulate in
//                       TMEM; two accumulator stages so the read-out of tile t
//                       overlaps the MMAs of tile t+1
//   warps 2-5 filter  : tcgen05.ld 32x32b -> thread = one corpus row, registers
//                       = scores against 32 queries; compare with tau[q];
//                       survivors are appended to the per-query candidate
//                       buffer (warp-aggregated atomics, rare after level 0)
//
// Roofline: HBM for small query batches (algorithmic bytes = rows * d * 2 per
// launch, streamed exactly once), tensor pipe for large ones (2 * rows * nq * d
// flops per launch).
#include "common.cuh"
#include <cstdio>

namespace kirag {

constexpr int kScanThreads = 192;
constexpr int kMaxStages = 12;
constexpr int kBlockBytes = kTileRows * 128;  // one [128 x 64] bf16 block = 16 KB
constexpr int kSmemLimit = 227 * 1024;

// ------------------------------------------------------------------ PTX ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must not hang the GPU — trap after ~4 s.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 8000000000LL) {
            printf("kirag scan_tc: mbarrier wait timed out (tag %d, block %d, thread %d, parity %u)\n", tag,
                   blockIdx.x, threadIdx.x, parity);
            __trap();
        }
    }
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_normal() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                         uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* holder_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(holder_smem)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// SWIZZLE_128B, K-major shared-memory matrix descriptor (sm_100 format):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (unused for swizzled K-major, 1)
//   [32,46) SBO >> 4 = 1024 B between 8-row groups | [46,48) version = 1 | [61,64) layout = 2
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major, N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct ScanArgs {
    const uint8_t* shadow;   // corpus bf16 blocks
    const uint8_t* qshadow;  // query bf16 blocks, tiles of BQ rows
    int64_t n_rows;
    int d;
    int64_t nq;
    int64_t tile_lo, tile_hi, n_tiles, tile_mult;
    const float* tau;  // [round_up(nq, 256)], +inf in the pad
    Cand* cand;        // [nq, cap]
    int* cnt;          // [nq]
    int cap;
    int n_stages;
    float* dump;       // optional [n_rows, dump_ld] dense approx scores (debug / tests)
    int64_t dump_ld;
};

// ---------------------------------------------------------------- kernel ----
template <int BQ, bool RESIDENT>
__global__ void __launch_bounds__(kScanThreads, 1) scan_tc_kernel(const ScanArgs a) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment for the 128B swizzle atoms
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int KC = a.d >> 6;                           // 64-wide k-blocks per row
    constexpr int kQBlockBytes = BQ * 128;             // one [BQ x 64] bf16 block
    const int NS = a.n_stages;
    const int stage_bytes = kBlockBytes + (RESIDENT ? 0 : kQBlockBytes);
    uint8_t* stage_base = smem;
    uint8_t* q_res = smem + (size_t)NS * stage_bytes;  // resident query tile (RESIDENT only)
    uint8_t* tail = q_res + (RESIDENT ? (size_t)KC * kQBlockBytes : 0);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
    uint64_t* empty_bar = full_bar + kMaxStages;
    uint64_t* tmem_full = empty_bar + kMaxStages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* q_bar = tmem_empty + 2;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(q_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    constexpr uint32_t kTmemCols = (2 * BQ <= 32) ? 32 : (2 * BQ <= 64) ? 64 : (2 * BQ <= 128) ? 128 : (2 * BQ <= 256) ? 256 : 512;
    const int n_qt = (int)((a.nq + BQ - 1) / BQ);

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 4); }
        mbar_init(q_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_holder, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0) {
        // =============================== producer ===============================
        if (lane == 0) {
            const uint64_t pol_x = (n_qt == 1) ? policy_evict_first() : policy_evict_normal();
            const uint64_t pol_q = policy_evict_last();
            if (RESIDENT) {
                mbar_expect_tx(q_bar, (uint32_t)(KC * kQBlockBytes));
                bulk_g2s(q_res, a.qshadow, (uint32_t)(KC * kQBlockBytes), q_bar, pol_q);
            }
            int s = 0;
            uint32_t ph = 0;
            for (int64_t ti = a.tile_lo + blockIdx.x; ti < a.tile_hi; ti += gridDim.x) {
                const int64_t tile = (ti * a.tile_mult) % a.n_tiles;
                const uint8_t* xsrc = a.shadow + (size_t)tile * ((size_t)a.d * kTileRows * 2);
                for (int qt = 0; qt < n_qt; ++qt) {
                    const uint8_t* qsrc = a.qshadow + (size_t)qt * ((size_t)a.d * BQ * 2);
                    for (int kc = 0; kc < KC; ++kc) {
                        mbar_wait(&empty_bar[s], ph ^ 1u, 100 + s);
                        uint8_t* dst = stage_base + (size_t)s * stage_bytes;
                        mbar_expect_tx(&full_bar[s], (uint32_t)stage_bytes);
                        bulk_g2s(dst, xsrc + (size_t)kc * kBlockBytes, kBlockBytes, &full_bar[s], pol_x);
                        if (!RESIDENT)
                            bulk_g2s(dst + kBlockBytes, qsrc + (size_t)kc * kQBlockBytes, kQBlockBytes, &full_bar[s], pol_q);
                        if (++s == NS) { s = 0; ph ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================= MMA ==================================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(kTileRows, BQ);
            if (RESIDENT) mbar_wait(q_bar, 0, 200);
            int s = 0;
            uint32_t ph = 0;
            uint32_t it = 0;
            for (int64_t ti = a.tile_lo + blockIdx.x; ti < a.tile_hi; ti += gridDim.x) {
                for (int qt = 0; qt < n_qt; ++qt, ++it) {
                    const uint32_t as = it & 1u;
                    const uint32_t aph = (it >> 1) & 1u;
                    mbar_wait(&tmem_empty[as], aph ^ 1u, 300 + as);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + as * BQ;
                    for (int kc = 0; kc < KC; ++kc) {
                        mbar_wait(&full_bar[s], ph, 400 + s);
                        tc_fence_after();
                        const uint32_t xa = smem_u32(stage_base + (size_t)s * stage_bytes);
                        const uint32_t qa = RESIDENT ? smem_u32(q_res + (size_t)kc * kQBlockBytes) : xa + kBlockBytes;
                        const uint64_t da = make_sw128_desc(xa);
                        const uint64_t db = make_sw128_desc(qa);
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4) {
                            // +32 bytes (two 16-byte units) per K=16 step inside the 128-byte swizzle row
                            umma_bf16(d_tmem, da + (uint64_t)(k4 * 2), db + (uint64_t)(k4 * 2), idesc,
                                      (kc | k4) ? 1u : 0u);
                        }
                        umma_commit(&empty_bar[s]);  // frees the stage once these MMAs have read it
                        if (++s == NS) { s = 0; ph ^= 1u; }
                    }
                    umma_commit(&tmem_full[as]);  // accumulator complete
                }
            }
        }
    } else {
        // ================================ filter ================================
        const int quad = warp & 3;  // TMEM lane quadrant this warp may read
        const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
        uint32_t it = 0;
        for (int64_t ti = a.tile_lo + blockIdx.x; ti < a.tile_hi; ti += gridDim.x) {
            const int64_t tile = (ti * a.tile_mult) % a.n_tiles;
            const int64_t row = tile * kTileRows + quad * 32 + lane;
            const bool row_ok = row < a.n_rows;
            for (int qt = 0; qt < n_qt; ++qt, ++it) {
                const uint32_t as = it & 1u;
                const uint32_t aph = (it >> 1) & 1u;
                mbar_wait(&tmem_full[as], aph, 500 + as);
                tc_fence_after();
#pragma unroll 1
                for (int g = 0; g < BQ / 32; ++g) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + lane_base + as * BQ + g * 32, v);
                    tmem_ld_wait();
                    if (g == BQ / 32 - 1) {
                        // all of this warp's reads of the accumulator are done
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tmem_empty[as]);
                    }
                    const int64_t q0 = (int64_t)qt * BQ + g * 32;
                    const float4* tp = reinterpret_cast<const float4*>(a.tau + q0);
                    uint32_t pass = 0;
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4) {
                        const float4 t = __ldg(tp + c4);
                        pass |= (__uint_as_float(v[c4 * 4 + 0]) >= t.x ? 1u : 0u) << (c4 * 4 + 0);
                        pass |= (__uint_as_float(v[c4 * 4 + 1]) >= t.y ? 1u : 0u) << (c4 * 4 + 1);
                        pass |= (__uint_as_float(v[c4 * 4 + 2]) >= t.z ? 1u : 0u) << (c4 * 4 + 2);
                        pass |= (__uint_as_float(v[c4 * 4 + 3]) >= t.w ? 1u : 0u) << (c4 * 4 + 3);
                    }
                    if (!row_ok) pass = 0;
                    if (a.dump && row_ok) {
#pragma unroll
                        for (int c = 0; c < 32; ++c)
                            if (q0 + c < a.nq) a.dump[row * a.dump_ld + q0 + c] = __uint_as_float(v[c]);
                    }
                    if (__any_sync(0xffffffffu, pass != 0)) {
                        // rare after the first level: warp-aggregated append, one atomic per (warp, query)
#pragma unroll
                        for (int c = 0; c < 32; ++c) {
                            const unsigned bal = __ballot_sync(0xffffffffu, (pass >> c) & 1u);
                            if (bal != 0 && q0 + c < a.nq) {
                                const int leader = __ffs(bal) - 1;
                                int base = 0;
                                if (lane == leader) base = atomicAdd(a.cnt + q0 + c, __popc(bal));
                                base = __shfl_sync(0xffffffffu, base, leader);
                                if ((pass >> c) & 1u) {
                                    const int slot = base + __popc(bal & ((1u << lane) - 1u));
                                    if (slot < a.cap) {
                                        Cand cd;
                                        cd.s = __uint_as_float(v[c]);
                                        cd.id = (int32_t)row;
                                        a.cand[(q0 + c) * (int64_t)a.cap + slot] = cd;
                                    }
                                }
                            }
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ------------------------------------------------------------------ host ----
int scan_tc_supported(int d) { return (d % 64 == 0 && d >= 64 && d <= 4096) ? 1 : 0; }

static size_t scan_smem_bytes(int bq, bool resident, int d, int n_stages) {
    const size_t stage = kBlockBytes + (resident ? 0 : (size_t)bq * 128);
    size_t total = 1024 /* alignment slack */ + (size_t)n_stages * stage;
    if (resident) total += (size_t)(d / 64) * bq * 128;
    total += (2 * kMaxStages + 5) * 8 + 16;
    return total;
}

static int pick_stages(int bq, bool resident, int d) {
    int ns = kMaxStages;
    while (ns > 0 && scan_smem_bytes(bq, resident, d, ns) > (size_t)kSmemLimit) --ns;
    return ns;
}

int scan_tc_pick(int64_t nq, int d, ScanTcPlan* plan) {
    KIRAG_CHECK(scan_tc_supported(d), "scan_tc: dimension %d is not supported (need a multiple of 64, <= 4096)", d);
    if (nq <= 32 && pick_stages(32, true, d) >= 4) { plan->bq = 32; plan->resident = 1; }
    else if (nq <= 64) { plan->bq = 64; plan->resident = 0; }
    else if (nq <= 128) { plan->bq = 128; plan->resident = 0; }
    else { plan->bq = 256; plan->resident = 0; }
    return 0;
}

size_t scan_tc_qshadow_bytes(int64_t nq, int d, const ScanTcPlan& plan) {
    const int64_t tiles = (nq + plan.bq - 1) / plan.bq;
    return (size_t)tiles * plan.bq * d * 2;
}

template <int BQ, bool RESIDENT>
static int launch_scan_t(const ScanArgs& args_in, int num_sms, cudaStream_t st) {
    ScanArgs args = args_in;
    args.n_stages = pick_stages(BQ, RESIDENT, args.d);
    KIRAG_CHECK(args.n_stages >= 2, "scan_tc: d=%d leaves no room for a shared-memory pipeline", args.d);
    const size_t smem = scan_smem_bytes(BQ, RESIDENT, args.d, args.n_stages);
    KIRAG_CUDA_OK(cudaFuncSetAttribute(scan_tc_kernel<BQ, RESIDENT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kSmemLimit));
    int64_t grid = args.tile_hi - args.tile_lo;
    if (grid > num_sms) grid = num_sms;
    if (grid <= 0) return 0;
    scan_tc_kernel<BQ, RESIDENT><<<(unsigned)grid, kScanThreads, smem, st>>>(args);
    KIRAG_LAUNCH_OK("scan_tc_kernel");
    return 0;
}

static int launch_scan_args(const ScanArgs& args, const ScanTcPlan& plan, int num_sms, cudaStream_t st) {
    if (plan.bq == 32 && plan.resident) return launch_scan_t<32, true>(args, num_sms, st);
    if (plan.bq == 64 && !plan.resident) return launch_scan_t<64, false>(args, num_sms, st);
    if (plan.bq == 128 && !plan.resident) return launch_scan_t<128, false>(args, num_sms, st);
    if (plan.bq == 256 && !plan.resident) return launch_scan_t<256, false>(args, num_sms, st);
    set_error("scan_tc: no kernel for plan (bq=%d resident=%d)", plan.bq, plan.resident);
    return 1;
}

int launch_scan_tc(const void* shadow, int64_t n_rows, int d, const void* qshadow, int64_t nq,
                   const ScanTcPlan& plan, int64_t tile_lo, int64_t tile_hi, int64_t n_tiles,
                   int64_t tile_mult, const float* tau, Cand* cand, int* cnt, int cap, int num_sms,
                   cudaStream_t st) {
    ScanArgs args{};
    args.shadow = (const uint8_t*)shadow;
    args.qshadow = (const uint8_t*)qshadow;
    args.n_rows = n_rows;
    args.d = d;
    args.nq = nq;
    args.tile_lo = tile_lo;
    args.tile_hi = tile_hi;
    args.n_tiles = n_tiles;
    args.tile_mult = tile_mult;
    args.tau = tau;
    args.cand = cand;
    args.cnt = cnt;
    args.cap = cap;
    args.dump = nullptr;
    args.dump_ld = 0;
    return launch_scan_args(args, plan, num_sms, st);
}

int launch_scan_tc_dump(const void* shadow, int64_t n_rows, int d, const void* qshadow, int64_t nq,
                        const ScanTcPlan& plan, const float* tau_inf, int* cnt_scratch, float* dump,
                        int64_t dump_ld, int num_sms, cudaStream_t st) {
    ScanArgs args{};
    args.shadow = (const uint8_t*)shadow;
    args.qshadow = (const uint8_t*)qshadow;
    args.n_rows = n_rows;
    args.d = d;
    args.nq = nq;
    args.n_tiles = (n_rows + kTileRows - 1) / kTileRows;
    args.tile_lo = 0;
    args.tile_hi = args.n_tiles;
    args.tile_mult = 1;
    args.tau = tau_inf;  // +inf everywhere: nothing is appended
    args.cand = nullptr;
    args.cnt = cnt_scratch;
    args.cap = 0;
    args.dump = dump;
    args.dump_ld = dump_ld;
    return launch_scan_args(args, plan, num_sms, st);
}

}  // namespace kirag
