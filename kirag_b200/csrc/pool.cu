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
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace kirag {

constexpr int kPoolCluster = 8;
constexpr int kPoolThreads = 256;
constexpr int kPoolMaxColsPerThread = 16;  // H <= 256 * 16 = 4096

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

// grid = B * kPoolCluster CTAs, cluster (kPoolCluster,1,1): cluster b <-> batch row b.
// dynamic smem: int tokens[S] | float partial[Hpad] | float red[kPoolThreads/32 + 2]
template <typename T, bool VEC>
__global__ void __cluster_dims__(kPoolCluster, 1, 1) __launch_bounds__(kPoolThreads)
pool_normalize_kernel(const T* __restrict__ hidden, const void* __restrict__ mask,
                      float* __restrict__ out, float* __restrict__ pooled_norm, int64_t S, int64_t H,
                      int64_t sb, int64_t ss, int64_t mb, int mask_dtype, int mode, int normalize) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int64_t b = blockIdx.x / kPoolCluster;
    extern __shared__ __align__(16) unsigned char pool_smem[];
    const int64_t Hpad = (H + 3) & ~(int64_t)3;
    int* tokens = reinterpret_cast<int*>(pool_smem);
    float* partial = reinterpret_cast<float*>(pool_smem + ((S * 4 + 15) & ~(int64_t)15));
    float* red = partial + Hpad;
    __shared__ int s_ntok;
    __shared__ float s_denom;

    // 1. every CTA of the cluster scans the mask row: kept-token list + denominator.
    //    (mean: denominator is the SUM of the mask values, kept = mask != 0, exactly
    //     like masked_fill(~mask.bool()) / mask.sum(1); cls: token 0, denominator 1)
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        int ntok = 0;
        long long msum = 0;
        if (mode == 1) {
            if (lane == 0) tokens[0] = 0;
            ntok = 1;
            msum = 1;
        } else {
            for (int64_t s0 = 0; s0 < S; s0 += 32) {
                const int64_t s = s0 + lane;
                const long long m = (s < S) ? load_mask(mask, mask_dtype, b * mb + s) : 0;
                msum += m;
                const unsigned bal = __ballot_sync(0xffffffffu, m != 0);
                if (m != 0) tokens[ntok + __popc(bal & ((1u << lane) - 1u))] = (int)s;
                ntok += __popc(bal);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) msum += __shfl_xor_sync(0xffffffffu, msum, o);
        }
        if (lane == 0) { s_ntok = ntok; s_denom = (float)msum; }
    }
    __syncthreads();
    const int ntok = s_ntok;

    // 2. this CTA sums tokens rank, rank+8, ... ; a thread owns 4 adjacent columns per pass
    const T* base = hidden + b * sb;
    float4 acc[kPoolMaxColsPerThread / 4];
#pragma unroll
    for (int i = 0; i < kPoolMaxColsPerThread / 4; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int n_pass = (int)((Hpad / 4 + kPoolThreads - 1) / kPoolThreads);
    // 4 tokens per trip so that 4 * n_pass independent 16-byte loads are in flight per thread
    for (int t0 = rank; t0 < ntok; t0 += 4 * kPoolCluster) {
        float4 v[4][kPoolMaxColsPerThread / 4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int t = t0 + u * kPoolCluster;
            const T* row = base + (int64_t)tokens[t < ntok ? t : t0] * ss;
#pragma unroll
            for (int p = 0; p < kPoolMaxColsPerThread / 4; ++p) {
                v[u][p] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p < n_pass && t < ntok) {
                    const int64_t c = ((int64_t)p * kPoolThreads + threadIdx.x) * 4;
                    if (VEC) {
                        if (c < H) v[u][p] = HiddenLoad<T>::load4(row + c);
                    } else {
                        if (c + 0 < H) v[u][p].x = HiddenLoad<T>::load1(row + c + 0);
                        if (c + 1 < H) v[u][p].y = HiddenLoad<T>::load1(row + c + 1);
                        if (c + 2 < H) v[u][p].z = HiddenLoad<T>::load1(row + c + 2);
                        if (c + 3 < H) v[u][p].w = HiddenLoad<T>::load1(row + c + 3);
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int p = 0; p < kPoolMaxColsPerThread / 4; ++p) {
                acc[p].x += v[u][p].x; acc[p].y += v[u][p].y;
                acc[p].z += v[u][p].z; acc[p].w += v[u][p].w;
            }
        }
    }
#pragma unroll
    for (int p = 0; p < kPoolMaxColsPerThread / 4; ++p) {
        if (p < n_pass) {
            const int64_t c = ((int64_t)p * kPoolThreads + threadIdx.x) * 4;
            if (c < Hpad) *reinterpret_cast<float4*>(partial + c) = acc[p];
        }
    }
    cluster.sync();

    // 3. CTA `rank` reduces its slice of columns over the 8 partials (DSMEM reads),
    //    divides by the denominator, and accumulates its share of ||p||^2
    const int64_t cols_per_rank = ((Hpad / 4 + kPoolCluster - 1) / kPoolCluster) * 4;
    const int64_t c_lo = rank * cols_per_rank;
    const int64_t c_hi = (c_lo + cols_per_rank < H) ? (c_lo + cols_per_rank) : H;
    const float denom = s_denom;
    float ss_local = 0.0f;
    for (int64_t c = c_lo + threadIdx.x; c < c_hi; c += kPoolThreads) {
        float sum = 0.0f;
#pragma unroll
        for (int r = 0; r < kPoolCluster; ++r) {
            const float* peer = cluster.map_shared_rank(partial, r);
            sum += peer[c];
        }
        const float pooled = sum / denom;
        ss_local = fmaf(pooled, pooled, ss_local);
        // stash the pooled value in the (now consumed) local column slot of the output
        out[b * H + c] = pooled;
    }
    ss_local = warp_butterfly_sum(ss_local);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss_local;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < kPoolThreads / 32; ++w) t += red[w];
        red[kPoolThreads / 32] = t;  // this CTA's share of the squared norm
    }
    cluster.sync();

    // 4. total norm from the 8 shares, normalise this CTA's slice
    float total = 0.f;
#pragma unroll
    for (int r = 0; r < kPoolCluster; ++r) {
        const float* peer = cluster.map_shared_rank(red, r);
        total += peer[kPoolThreads / 32];
    }
    const float nrm = sqrtf(total);
    if (rank == 0 && threadIdx.x == 0 && pooled_norm) pooled_norm[b] = nrm;
    if (normalize) {
        const float dn = fmaxf(nrm, 1e-12f);
        for (int64_t c = c_lo + threadIdx.x; c < c_hi; c += kPoolThreads)
            out[b * H + c] = out[b * H + c] / dn;
    }
    // peers may still be reading this CTA's shared memory
    cluster.sync();
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

template <typename T>
static int pool_launch_t(const void* hidden, const void* mask, float* out, float* pooled_norm,
                         int64_t B, int64_t S, int64_t H, int64_t sb, int64_t ss, int64_t mb,
                         int mask_dtype, int mode, int normalize, cudaStream_t st) {
    const int64_t Hpad = (H + 3) & ~(int64_t)3;
    const size_t smem = ((S * 4 + 15) & ~(size_t)15) + (size_t)Hpad * 4 + (kPoolThreads / 32 + 4) * 4;
    KIRAG_CHECK(smem <= 200 * 1024, "pool_normalize: S=%lld H=%lld need %zu B of shared memory",
                (long long)S, (long long)H, smem);
    const int vec_elems = 4;
    const size_t esz = sizeof(T);
    const bool vec = (H % vec_elems == 0) && (sb % vec_elems == 0) && (ss % vec_elems == 0) &&
                     ((reinterpret_cast<uintptr_t>(hidden) % (esz * vec_elems)) == 0);
    const unsigned grid = (unsigned)(B * kPoolCluster);
    if (vec) {
        if (smem > 48 * 1024 && ensure_dynamic_smem(pool_normalize_kernel<T, true>, smem)) return 1;
        pool_normalize_kernel<T, true><<<grid, kPoolThreads, smem, st>>>(
            (const T*)hidden, mask, out, pooled_norm, S, H, sb, ss, mb, mask_dtype, mode, normalize);
    } else {
        if (smem > 48 * 1024 && ensure_dynamic_smem(pool_normalize_kernel<T, false>, smem)) return 1;
        pool_normalize_kernel<T, false><<<grid, kPoolThreads, smem, st>>>(
            (const T*)hidden, mask, out, pooled_norm, S, H, sb, ss, mb, mask_dtype, mode, normalize);
    }
    KIRAG_LAUNCH_OK("pool_normalize_kernel");
    return 0;
}

int launch_pool_normalize(const void* hidden, const void* mask, float* out, float* pooled_norm,
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
        case 0: return pool_launch_t<float>(hidden, mask, out, pooled_norm, B, S, H, sb, ss, mb, mask_dtype, mode, normalize, st);
        case 1: return pool_launch_t<__nv_bfloat16>(hidden, mask, out, pooled_norm, B, S, H, sb, ss, mb, mask_dtype, mode, normalize, st);
        case 2: return pool_launch_t<__half>(hidden, mask, out, pooled_norm, B, S, H, sb, ss, mb, mask_dtype, mode, normalize, st);
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
