// pool.cu — masked mean-pool (or CLS-select) + L2-normalise embedding epilogue.
//
// Replaces, in one pass over the hidden states:
//   average_pool            /root/reference/retriever/encoders.py:56-58  (dup. retriever/e5.py:46-48)
//       last_hidden.masked_fill(~mask[..., None].bool(), 0).sum(1) / mask.sum(1)[..., None]
//   F.normalize(p=2, dim=1) /root/reference/retriever/encoders.py:76     (dup. retriever/e5.py:60)
//       x / max(||x||_2, 1e-12)
//   BGE tail                /root/reference/retriever/encoders.py:115-117
//       normalize(last_hidden[:, 0])
// The reference runs 5-7 ATen launches and ~3 passes over [B,S,H]; here a
// cluster of 8 CTAs owns one batch row, each CTA streams whole 4*H-byte token
// rows (only the unmasked ones), the 8 partial sums meet in distributed
// shared memory, and the row is normalised and written once.
//
// HBM-bound.  Algorithmic bytes per launch = sum_b len_b * H * sizeof(hidden)
// + B*S*sizeof(mask) + B*H*4.
#include "common.cuh"
#include <cstdlib>
#include <cmath>
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace kirag {

constexpr int kPoolCluster = 8;
constexpr int kPoolThreads = 256;
constexpr int kPoolMaxColsPerThread = 16;  // H <= 256 * 16 = 4096 (limit checked by the launcher)

template <typename T> struct HiddenLoad;
template <> struct HiddenLoad<float> {
    static __device__ __forceinline__ float4 load4(const float* p) {
        return __ldcs(reinterpret_cast<const float4*>(p));
    }
    static __device__ __forceinline__ float load1(const float* p) { return p[0]; }
    static __device__ __forceinline__ void store4(float* p, float4 v) {
        *reinterpret_cast<float4*>(p) = v;
    }
    static __device__ __forceinline__ void store1(float* p, float v) { p[0] = v; }
};
template <> struct HiddenLoad<__nv_bfloat16> {
    static __device__ __forceinline__ float4 load4(const __nv_bfloat16* p) {
        const uint2 raw = __ldcs(reinterpret_cast<const uint2*>(p));
        const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&raw.x);
        const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
        const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
        return make_float4(fa.x, fa.y, fb.x, fb.y);
    }
    static __device__ __forceinline__ float load1(const __nv_bfloat16* p) {
        return __bfloat162float(p[0]);
    }
    static __device__ __forceinline__ void store4(__nv_bfloat16* p, float4 v) {
        __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        uint2 raw;
        raw.x = *reinterpret_cast<uint32_t*>(&a);
        raw.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(p) = raw;
    }
    static __device__ __forceinline__ void store1(__nv_bfloat16* p, float v) {
        p[0] = __float2bfloat16_rn(v);
    }
};
template <> struct HiddenLoad<__half> {
    static __device__ __forceinline__ float4 load4(const __half* p) {
        const uint2 raw = __ldcs(reinterpret_cast<const uint2*>(p));
        const __half2 a = *reinterpret_cast<const __half2*>(&raw.x);
        const __half2 b = *reinterpret_cast<const __half2*>(&raw.y);
        const float2 fa = __half22float2(a), fb = __half22float2(b);
        return make_float4(fa.x, fa.y, fb.x, fb.y);
    }
    static __device__ __forceinline__ float load1(const __half* p) { return __half2float(p[0]); }
    static __device__ __forceinline__ void store4(__half* p, float4 v) {
        __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
        uint2 raw;
        raw.x = *reinterpret_cast<uint32_t*>(&a);
        raw.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(p) = raw;
    }
    static __device__ __forceinline__ void store1(__half* p, float v) { p[0] = __float2half_rn(v); }
};

__device__ __forceinline__ long long load_mask(const void* mask, int mask_dtype, int64_t idx) {
    if (mask_dtype == 0) return reinterpret_cast<const long long*>(mask)[idx];
    if (mask_dtype == 1) return reinterpret_cast<const int*>(mask)[idx];
    return reinterpret_cast<const unsigned char*>(mask)[idx];
}

// 16-byte loads of the hidden states, whatever the dtype (4 fp32 / 8 bf16 / 8 fp16 columns).
// The raw 16 bytes stay in 4 registers until they are accumulated, so that 8 independent loads
// per thread are in flight at a modest register cost.
template <typename T> struct Vec16;
template <> struct Vec16<float> {
    static constexpr int kElems = 4;
    static __device__ __forceinline__ void add(const uint4& raw, float scale, float (&acc)[8]) {
        acc[0] = fmaf(__uint_as_float(raw.x), scale, acc[0]);
        acc[1] = fmaf(__uint_as_float(raw.y), scale, acc[1]);
        acc[2] = fmaf(__uint_as_float(raw.z), scale, acc[2]);
        acc[3] = fmaf(__uint_as_float(raw.w), scale, acc[3]);
    }
};
template <> struct Vec16<__nv_bfloat16> {
    static constexpr int kElems = 8;
    static __device__ __forceinline__ void add(const uint4& raw, float scale, float (&acc)[8]) {
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            // bf16 -> fp32 is a 16-bit shift
            acc[2 * i] = fmaf(__uint_as_float(w[i] << 16), scale, acc[2 * i]);
            acc[2 * i + 1] = fmaf(__uint_as_float(w[i] & 0xffff0000u), scale, acc[2 * i + 1]);
        }
    }
};
template <> struct Vec16<__half> {
    static constexpr int kElems = 8;
    static __device__ __forceinline__ void add(const uint4& raw, float scale, float (&acc)[8]) {
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
            acc[2 * i] = fmaf(t.x, scale, acc[2 * i]);
            acc[2 * i + 1] = fmaf(t.y, scale, acc[2 * i + 1]);
        }
    }
};

// Cluster of CL CTAs (CL = 1, 2, 4 or 8, chosen at launch so that the grid fills the GPU) owns
// one batch row.  Shared memory: tokens[S] | ballots[ceil(S/32)] | prefix[ceil(S/32)+1] |
// partial[R][Hpad] | red[16]
//   1. all 256 threads scan the mask row in parallel (one ballot per 32 tokens, one warp-level
//      scan of the ballot popcounts) -> list of kept tokens + denominator (= SUM of mask values,
//      kept = mask != 0, exactly masked_fill(~mask.bool()) / mask.sum(1); cls: token 0, denom 1)
//   2. CTA `rank` streams kept tokens rank, rank+CL, ...: the CTA's (token, 16-byte chunk) pairs
//      are dealt to the threads in order, so every thread issues full 16-byte loads whatever the
//      dtype and row length; 8 independent loads in flight per thread
//   3. the partial sums meet in distributed shared memory: CTA `rank` reduces its slice of the
//      columns over all ranks (and sub-rows), divides, accumulates ||p||^2; a second DSMEM
//      exchange of the CL shares gives the norm; each CTA normalises and writes its slice.
// VEC path requires H % kElems == 0 and (256 % TPR == 0 or TPR % 256 == 0), TPR = H / kElems.
constexpr int kPoolUnroll = 8;
constexpr int kPoolMaxAcc = 4;  // TPR <= 1024 chunks per row

template <typename T, bool VEC, int NACC>
__global__ void __launch_bounds__(kPoolThreads)
pool_normalize_kernel(const T* __restrict__ hidden, const void* __restrict__ mask,
                      float* __restrict__ out, T* __restrict__ out_lp, float* __restrict__ pooled_norm, int64_t S, int64_t H,
                      int64_t sb, int64_t ss, int64_t mb, int mask_dtype, int mode, int normalize) {
    cg::cluster_group cluster = cg::this_cluster();
    const int CL = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int64_t b = blockIdx.x / CL;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int E = Vec16<T>::kElems;
    const int64_t Hpad = (H + 7) & ~(int64_t)7;
    const int n_words = (int)((S + 31) >> 5);
    const int TPR = VEC ? (int)(H / E) : (int)H;               // 16-byte chunks (or scalars) per token row
    const int R = VEC ? (TPR >= kPoolThreads ? 1 : kPoolThreads / TPR) : 1;  // token rows per CTA-wide load
    extern __shared__ __align__(16) unsigned char pool_smem[];
    int* tokens = reinterpret_cast<int*>(pool_smem);
    unsigned* ballots = reinterpret_cast<unsigned*>(tokens + S);
    int* prefix = reinterpret_cast<int*>(ballots + n_words);
    float* partial = reinterpret_cast<float*>(pool_smem + ((((size_t)S + 2 * (size_t)n_words + 1) * 4 + 15) & ~(size_t)15));
    float* red = partial + (size_t)R * Hpad;
    __shared__ int s_ntok;
    __shared__ float s_denom;
    __shared__ unsigned long long s_msum;

    // ---- 1. kept-token list ------------------------------------------------------------------
    if (mode == 1) {
        if (tid == 0) { tokens[0] = 0; s_ntok = 1; s_denom = 1.0f; }
        __syncthreads();
    } else {
        if (tid == 0) s_msum = 0ull;
        __syncthreads();
        long long msum = 0;
        for (int64_t s0 = 0; s0 < S; s0 += kPoolThreads) {
            const int64_t sidx = s0 + tid;
            const long long mv = (sidx < S) ? load_mask(mask, mask_dtype, b * mb + sidx) : 0;
            msum += mv;
            const unsigned bal = __ballot_sync(0xffffffffu, mv != 0);
            if (lane == 0 && (s0 >> 5) + warp < n_words) ballots[(s0 >> 5) + warp] = bal;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) msum += __shfl_xor_sync(0xffffffffu, msum, o);
        if (lane == 0) atomicAdd(&s_msum, (unsigned long long)msum);
        __syncthreads();
        if (warp == 0) {
            int carry = 0;
            for (int w0 = 0; w0 < n_words; w0 += 32) {
                const int w = w0 + lane;
                const int c = (w < n_words) ? __popc(ballots[w]) : 0;
                int incl = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                if (w < n_words) prefix[w] = carry + incl - c;
                carry += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (lane == 0) { s_ntok = carry; s_denom = (float)(long long)s_msum; }
        }
        __syncthreads();
        for (int64_t s0 = 0; s0 < S; s0 += kPoolThreads) {
            const int64_t sidx = s0 + tid;
            if (sidx < S) {
                const unsigned bal = ballots[sidx >> 5];
                if ((bal >> lane) & 1u) tokens[prefix[sidx >> 5] + __popc(bal & ((1u << lane) - 1u))] = (int)sidx;
            }
        }
        __syncthreads();
    }
    const int ntok = s_ntok;

    // ---- 2. stream this CTA's tokens ---------------------------------------------------------
    const T* base = hidden + b * sb;
    const int my_ntok = (ntok > rank) ? (ntok - rank + CL - 1) / CL : 0;  // tokens rank, rank+CL, ...
    if (VEC) {
        // thread's fixed position inside a token row: chunk(s) c_a = c0 + 256*a, sub-row r (TPR < 256)
        const int c0 = (TPR >= kPoolThreads) ? tid : tid % TPR;
        const int r = (TPR >= kPoolThreads) ? 0 : tid / TPR;
        float acc[NACC][8];
#pragma unroll
        for (int a = 0; a < NACC; ++a)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[a][i] = 0.f;
        constexpr int U = kPoolUnroll / NACC;  // tokens in flight per thread (NACC loads each)
        for (int lt0 = r; lt0 < my_ntok; lt0 += R * U) {
            uint4 raw[U][NACC];
            float w[U];
            // all U*NACC loads are issued unconditionally (a slot past the end re-reads token lt0 with
            // weight 0), so nothing sits between them and they are all in flight together
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int lt = lt0 + u * R;
                const bool ok = lt < my_ntok;
                w[u] = ok ? 1.0f : 0.0f;
                const T* row = base + (int64_t)tokens[rank + (ok ? lt : lt0) * CL] * ss + (int64_t)c0 * E;
#pragma unroll
                for (int a = 0; a < NACC; ++a)
                    raw[u][a] = __ldcs(reinterpret_cast<const uint4*>(row + (int64_t)a * kPoolThreads * E));
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int a = 0; a < NACC; ++a) Vec16<T>::add(raw[u][a], w[u], acc[a]);
        }
#pragma unroll
        for (int a = 0; a < NACC; ++a) {
            float* dst = partial + (size_t)r * Hpad + (size_t)(c0 + a * kPoolThreads) * E;
#pragma unroll
            for (int i = 0; i < E; ++i) dst[i] = acc[a][i];
        }
    } else {
        // scalar fallback (odd H / unaligned views): thread owns columns tid, tid+256, ...
        for (int64_t c = tid; c < H; c += kPoolThreads) {
            float sum = 0.f;
            for (int t = rank; t < ntok; t += CL) sum += HiddenLoad<T>::load1(base + (int64_t)tokens[t] * ss + c);
            partial[c] = sum;
        }
    }
    cluster.sync();

    // ---- 3. cross-CTA reduction in distributed shared memory ---------------------------------
    const int64_t cols_per_rank = ((H + CL - 1) / CL + 3) & ~(int64_t)3;
    const int64_t c_lo = rank * cols_per_rank;
    const int64_t c_hi = (c_lo + cols_per_rank < H) ? (c_lo + cols_per_rank) : H;
    const float denom = s_denom;
    float ss_local = 0.0f;
    for (int64_t c = c_lo + tid; c < c_hi; c += kPoolThreads) {
        float sum = 0.0f;
        for (int r = 0; r < CL; ++r) {
            const float* peer = cluster.map_shared_rank(partial, r);
            for (int rr = 0; rr < R; ++rr) sum += peer[(size_t)rr * Hpad + c];
        }
        const float pooled = sum / denom;
        ss_local = fmaf(pooled, pooled, ss_local);
        out[b * H + c] = pooled;  // re-read below by the same thread when normalising
    }
    ss_local = warp_butterfly_sum(ss_local);
    if (lane == 0) red[warp] = ss_local;
    __syncthreads();
    if (tid == 0) {
        float t = 0.f;
        for (int w = 0; w < kPoolThreads / 32; ++w) t += red[w];
        red[kPoolThreads / 32] = t;  // this CTA's share of the squared norm
    }
    cluster.sync();
    float total_ss = 0.f;
    for (int r = 0; r < CL; ++r) {
        const float* peer = cluster.map_shared_rank(red, r);
        total_ss += peer[kPoolThreads / 32];
    }
    const float nrm = sqrtf(total_ss);
    if (rank == 0 && tid == 0 && pooled_norm) pooled_norm[b] = nrm;
    if (normalize || out_lp) {
        // out_lp (optional): the result once more in the dtype of the hidden states — what the reference's ops return
        // for bf16 / fp16 inputs — written here instead of by a separate cast kernel
        const float dn = normalize ? fmaxf(nrm, 1e-12f) : 1.0f;
        for (int64_t c = c_lo + tid; c < c_hi; c += kPoolThreads) {
            const float v = out[b * H + c] / dn;
            if (normalize) out[b * H + c] = v;
            if (out_lp) HiddenLoad<T>::store1(out_lp + b * H + c, v);
        }
    }
    cluster.sync();  // peers may still be reading this CTA's shared memory
}

// backward: grad_hidden[b,s,:] = keep(b,s)/denom_b * gp_b,
//   gp = g                          (normalize == 0)
//   gp = (g - y*(y.g)) / max(n,eps) (normalize == 1, n > eps), g/eps when n <= eps
// grid (B, s_chunks)
template <typename T>
__global__ void __launch_bounds__(256)
pool_normalize_backward_kernel(const float* __restrict__ grad_out, const float* __restrict__ out,
                               const float* __restrict__ pooled_norm, const void* __restrict__ mask,
                               T* __restrict__ grad_hidden, int64_t S, int64_t H, int64_t mb,
                               int mask_dtype, int mode, int normalize, int s_per_cta) {
    const int64_t b = blockIdx.x;
    __shared__ float red[10];
    __shared__ float s_denom;
    const float* g = grad_out + b * H;
    const float* y = out + b * H;
    float dot = 0.f;
    if (normalize) {
        for (int64_t c = threadIdx.x; c < H; c += blockDim.x) dot = fmaf(g[c], y[c], dot);
        dot = warp_butterfly_sum(dot);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
    }
    // denominator
    if (threadIdx.x >= 32 && threadIdx.x < 64) {
        const int lane = threadIdx.x - 32;
        long long msum = 0;
        if (mode == 1) msum = 1;
        else {
            for (int64_t s = lane; s < S; s += 32) msum += load_mask(mask, mask_dtype, b * mb + s);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) msum += __shfl_xor_sync(0xffffffffu, msum, o);
        }
        if (lane == 0) s_denom = (float)msum;
    }
    __syncthreads();
    if (normalize) {
        dot = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) dot += red[w];
    }
    const float denom = s_denom;
    const float nrm = normalize ? pooled_norm[b] : 1.f;
    const bool clamped = normalize && !(nrm > 1e-12f);
    const float inv_n = normalize ? 1.f / fmaxf(nrm, 1e-12f) : 1.f;
    const int64_t s_lo = (int64_t)blockIdx.y * s_per_cta;
    const int64_t s_hi = (s_lo + s_per_cta < S) ? (s_lo + s_per_cta) : S;
    for (int64_t s = s_lo; s < s_hi; ++s) {
        bool keep;
        if (mode == 1) keep = (s == 0);
        else keep = load_mask(mask, mask_dtype, b * mb + s) != 0;
        T* dst = grad_hidden + (b * S + s) * H;
        for (int64_t c = threadIdx.x; c < H; c += blockDim.x) {
            float v = 0.f;
            if (keep) {
                float gp = g[c];
                if (normalize) gp = clamped ? gp * inv_n : (gp - y[c] * dot) * inv_n;
                v = gp / denom;
            }
            HiddenLoad<T>::store1(dst + c, v);
        }
    }
}

template <typename T, bool VEC, int NACC>
static int pool_launch_cfg(const void* hidden, const void* mask, float* out, void* out_lp, float* pooled_norm, int64_t B, int64_t S,
                           int64_t H, int64_t sb, int64_t ss, int64_t mb, int mask_dtype, int mode, int normalize,
                           int cl, size_t smem, cudaStream_t st) {
    if (smem > 48 * 1024 && ensure_dynamic_smem(pool_normalize_kernel<T, VEC, NACC>, smem)) return 1;
    if (cl == 0) {
        // Cluster size (CTAs per batch row): of the sizes 1/2/4/8 that give at least 3 CTAs per SM's worth of grid,
        // the one whose grid fills its last wave best, the smallest on a tie.  Measured (tools/pool_ab.py, full masks,
        // [B,512,1024]): B=256 2 CTAs per row 85 us fp32 / 50 us bf16 vs 98 / 60 us with 4 (1.38 waves); B=1024 2 CTAs per
        // row 319 us vs 331 us with 1.
        static int per_sm_cache = 0;
        static size_t per_sm_smem = (size_t)-1;
        if (per_sm_smem != smem) {
            int n = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, pool_normalize_kernel<T, VEC, NACC>, kPoolThreads, smem) !=
                    cudaSuccess || n <= 0) { cudaGetLastError(); n = 4; }
            per_sm_cache = n;
            per_sm_smem = smem;
        }
        const double slots = 148.0 * per_sm_cache;
        double best = -1.0;
        cl = 8;
        for (int c = 1; c <= 8; c *= 2) {
            const double ctas = (double)B * c;
            if (ctas < 148.0 * 3 && c < 8) continue;
            const double waves = ctas / slots;
            const double eff = waves <= 1.0 ? 1.0 : waves / ceil(waves);  // one partial wave has no tail
            if (eff > best + 1e-9) { best = eff; cl = c; }
        }
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(B * cl));
    cfg.blockDim = dim3(kPoolThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const T* hp = (const T*)hidden;
    KIRAG_CUDA_OK(cudaLaunchKernelEx(&cfg, pool_normalize_kernel<T, VEC, NACC>, hp, mask, out, (T*)out_lp, pooled_norm, S, H, sb, ss, mb,
                                     mask_dtype, mode, normalize));
    count_launch();
    return 0;
}

template <typename T>
static int pool_launch_t(const void* hidden, const void* mask, float* out, void* out_lp, float* pooled_norm,
                         int64_t B, int64_t S, int64_t H, int64_t sb, int64_t ss, int64_t mb,
                         int mask_dtype, int mode, int normalize, cudaStream_t st) {
    constexpr int E = Vec16<T>::kElems;
    const int64_t Hpad = (H + 7) & ~(int64_t)7;
    const int64_t tpr = (H % E == 0) ? H / E : 0;
    const bool vec = tpr > 0 && tpr <= (int64_t)kPoolThreads * kPoolMaxAcc &&
                     ((kPoolThreads % tpr == 0) || (tpr % kPoolThreads == 0)) && (sb % E == 0) && (ss % E == 0) &&
                     ((reinterpret_cast<uintptr_t>(hidden) & 15) == 0);
    const int R = vec ? (tpr >= kPoolThreads ? 1 : (int)(kPoolThreads / tpr)) : 1;
    const int64_t n_words = (S + 31) / 32;
    const size_t smem = ((((size_t)S + 2 * (size_t)n_words + 1) * 4 + 15) & ~(size_t)15) + (size_t)R * Hpad * 4 + 16 * 4;
    KIRAG_CHECK(smem <= 200 * 1024, "pool_normalize: S=%lld H=%lld need %zu B of shared memory",
                (long long)S, (long long)H, smem);
    int cl = 0;  // 0: chosen per kernel instantiation in pool_launch_cfg (needs its occupancy)
    {   // experiment knob
        const char* v = getenv("KIRAG_POOL_CL");
        if (v && *v) { const int c = atoi(v); if (c == 1 || c == 2 || c == 4 || c == 8) cl = c; }
    }
    if (mode == 1) cl = 1;  // CLS: one token per row
#define KIRAG_POOL_GO(VECF, NACC) \
    pool_launch_cfg<T, VECF, NACC>(hidden, mask, out, out_lp, pooled_norm, B, S, H, sb, ss, mb, mask_dtype, mode, normalize, cl, smem, st)
    if (!vec) return KIRAG_POOL_GO(false, 1);
    const int nacc = tpr >= kPoolThreads ? (int)(tpr / kPoolThreads) : 1;
    if (nacc == 1) return KIRAG_POOL_GO(true, 1);
    if (nacc == 2) return KIRAG_POOL_GO(true, 2);
    if (nacc == 4) return KIRAG_POOL_GO(true, 4);
    return KIRAG_POOL_GO(false, 1);  // 3 chunks per thread etc.: scalar path
#undef KIRAG_POOL_GO
}

int launch_pool_normalize(const void* hidden, const void* mask, float* out, void* out_lp, float* pooled_norm,
                          int64_t B, int64_t S, int64_t H, int64_t sb, int64_t ss, int64_t mb,
                          int hidden_dtype, int mask_dtype, int mode, int normalize,
                          cudaStream_t st) {
    if (B <= 0) return 0;
    KIRAG_CHECK(H > 0 && S > 0, "pool_normalize: empty S=%lld or H=%lld", (long long)S, (long long)H);
    KIRAG_CHECK(H <= (int64_t)kPoolThreads * kPoolMaxColsPerThread,
                "pool_normalize: H=%lld exceeds %d", (long long)H, kPoolThreads * kPoolMaxColsPerThread);
    KIRAG_CHECK(B * kPoolCluster < 0x7fffffffLL, "pool_normalize: B=%lld too large", (long long)B);
    KIRAG_CHECK(mode == 0 || mode == 1, "pool_normalize: unknown mode %d", mode);
    KIRAG_CHECK(mask_dtype >= 0 && mask_dtype <= 2, "pool_normalize: unknown mask dtype %d", mask_dtype);
    KIRAG_CHECK(mode == 1 || mask != nullptr, "pool_normalize: mean mode needs a mask");
    switch (hidden_dtype) {
        case 0: return pool_launch_t<float>(hidden, mask, out, out_lp, pooled_norm, B, S, H, sb, ss, mb, mask_dtype, mode, normalize, st);
        case 1: return pool_launch_t<__nv_bfloat16>(hidden, mask, out, out_lp, pooled_norm, B, S, H, sb, ss, mb, mask_dtype, mode, normalize, st);
        case 2: return pool_launch_t<__half>(hidden, mask, out, out_lp, pooled_norm, B, S, H, sb, ss, mb, mask_dtype, mode, normalize, st);
        default: break;
    }
    set_error("pool_normalize: unknown hidden dtype %d", hidden_dtype);
    return 1;
}

int launch_pool_normalize_backward(const float* grad_out, const float* out, const float* pooled_norm,
                                   const void* mask, void* grad_hidden, int64_t B, int64_t S,
                                   int64_t H, int64_t mb, int hidden_dtype, int mask_dtype, int mode,
                                   int normalize, cudaStream_t st) {
    if (B <= 0) return 0;
    KIRAG_CHECK(mode == 0 || mode == 1, "pool_normalize_backward: unknown mode %d", mode);
    KIRAG_CHECK(B <= 0x7fffffffLL, "pool_normalize_backward: B too large");
    const int s_per_cta = 16;
    dim3 grid((unsigned)B, (unsigned)((S + s_per_cta - 1) / s_per_cta));
    KIRAG_CHECK(grid.y <= 65535, "pool_normalize_backward: S=%lld too large", (long long)S);
    switch (hidden_dtype) {
        case 0:
            pool_normalize_backward_kernel<float><<<grid, 256, 0, st>>>(
                grad_out, out, pooled_norm, mask, (float*)grad_hidden, S, H, mb, mask_dtype, mode, normalize, s_per_cta);
            break;
        case 1:
            pool_normalize_backward_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(
                grad_out, out, pooled_norm, mask, (__nv_bfloat16*)grad_hidden, S, H, mb, mask_dtype, mode, normalize, s_per_cta);
            break;
        case 2:
            pool_normalize_backward_kernel<__half><<<grid, 256, 0, st>>>(
                grad_out, out, pooled_norm, mask, (__half*)grad_hidden, S, H, mb, mask_dtype, mode, normalize, s_per_cta);
            break;
        default:
            set_error("pool_normalize_backward: unknown hidden dtype %d", hidden_dtype);
            return 1;
    }
    KIRAG_LAUNCH_OK("pool_normalize_backward_kernel");
    return 0;
}

}  // namespace kirag
