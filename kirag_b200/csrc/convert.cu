// convert.cu — fp32 rows -> bf16 shadow in the swizzled block layout, plus
// the row norms the exactness certificate needs.
//
// Replaces the storage half of faiss.IndexFlatIP.add (reference call site
// /root/reference/retriever/index.py:32): FAISS appends the fp32 rows to its
// code vector; this build keeps the same fp32 master and additionally writes
// a bf16 shadow that the tcgen05 filter scan streams.
#include "common.cuh"

namespace kirag {

// One warp per row.  Lane l handles 16-byte output granules l, l+32, ... (8
// columns each): reads 32 contiguous bytes of fp32, writes 16 bytes of bf16.
// A warp therefore reads 1 KB contiguous per step and fills whole 128-byte
// swizzle rows.  HBM-bound: 4 B read + 2 B written per element.
//
// `center` (optional, corpus rows only): the shadow stores bf16(x - c) for a fixed vector c.  Subtracting a
// constant vector from every corpus row shifts all scores of a query by the constant <q, c>, so the ranking the
// filter works on is unchanged, while the bf16 rounding error now scales with ||x - c|| instead of ||x||: on
// embeddings with a large common component (E5: random-pair cosine 0.7+) that is what keeps the exactness
// certificate passing (api.cu::cert_eps).  maxnorm2_bits: [0] max ||y||^2, [1] max ||y - bf16(y)||^2 with
// y = fl(x - c) (y = x without a centre), [2] max ||x||^2.
__global__ void __launch_bounds__(256)
convert_rows_kernel(const float* __restrict__ src, int64_t n_rows, int d, int64_t dst_row0,
                    uint8_t* __restrict__ shadow, int rows_per_tile,
                    unsigned* __restrict__ maxnorm2_bits, float* __restrict__ row_norms,
                    float* __restrict__ row_errs, const float* __restrict__ center, float* __restrict__ row_cdot) {
    pdl_wait();
    pdl_launch_dependents();
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int n_gran = d >> 3;
    for (int64_t r = warp; r < n_rows; r += n_warps) {
        const float* row = src + r * (int64_t)d;
        float ss = 0.0f;  // ||y||^2
        float xs = 0.0f;  // ||x||^2 (differs from ss only with a centre)
        float es = 0.0f;  // squared norm of the bf16 rounding error of this row (exact differences, fp32 sum)
        float cd = 0.0f;  // <x, c> (queries: the score offset the centred shadow leaves out)
        for (int g = lane; g < n_gran; g += 32) {
            float4 a = *reinterpret_cast<const float4*>(row + g * 8);
            float4 b = *reinterpret_cast<const float4*>(row + g * 8 + 4);
            if (center) {
                const float4 ca = __ldg(reinterpret_cast<const float4*>(center + g * 8));
                const float4 cb = __ldg(reinterpret_cast<const float4*>(center + g * 8 + 4));
                if (row_cdot) {  // query side: the row itself is NOT shifted, only <q, c> is wanted
                    cd = fmaf(a.x, ca.x, cd); cd = fmaf(a.y, ca.y, cd); cd = fmaf(a.z, ca.z, cd); cd = fmaf(a.w, ca.w, cd);
                    cd = fmaf(b.x, cb.x, cd); cd = fmaf(b.y, cb.y, cd); cd = fmaf(b.z, cb.z, cd); cd = fmaf(b.w, cb.w, cd);
                } else {
                    xs = fmaf(a.x, a.x, xs); xs = fmaf(a.y, a.y, xs); xs = fmaf(a.z, a.z, xs); xs = fmaf(a.w, a.w, xs);
                    xs = fmaf(b.x, b.x, xs); xs = fmaf(b.y, b.y, xs); xs = fmaf(b.z, b.z, xs); xs = fmaf(b.w, b.w, xs);
                    a.x = __fsub_rn(a.x, ca.x); a.y = __fsub_rn(a.y, ca.y); a.z = __fsub_rn(a.z, ca.z); a.w = __fsub_rn(a.w, ca.w);
                    b.x = __fsub_rn(b.x, cb.x); b.y = __fsub_rn(b.y, cb.y); b.z = __fsub_rn(b.z, cb.z); b.w = __fsub_rn(b.w, cb.w);
                }
            }
            ss = fmaf(a.x, a.x, ss); ss = fmaf(a.y, a.y, ss);
            ss = fmaf(a.z, a.z, ss); ss = fmaf(a.w, a.w, ss);
            ss = fmaf(b.x, b.x, ss); ss = fmaf(b.y, b.y, ss);
            ss = fmaf(b.z, b.z, ss); ss = fmaf(b.w, b.w, ss);
            __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y);
            __nv_bfloat162 p1 = __floats2bfloat162_rn(a.z, a.w);
            __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y);
            __nv_bfloat162 p3 = __floats2bfloat162_rn(b.z, b.w);
            {
                const float2 r0 = __bfloat1622float2(p0), r1 = __bfloat1622float2(p1);
                const float2 r2 = __bfloat1622float2(p2), r3 = __bfloat1622float2(p3);
                float e;
                e = a.x - r0.x; es = fmaf(e, e, es); e = a.y - r0.y; es = fmaf(e, e, es);
                e = a.z - r1.x; es = fmaf(e, e, es); e = a.w - r1.y; es = fmaf(e, e, es);
                e = b.x - r2.x; es = fmaf(e, e, es); e = b.y - r2.y; es = fmaf(e, e, es);
                e = b.z - r3.x; es = fmaf(e, e, es); e = b.w - r3.y; es = fmaf(e, e, es);
            }
            uint4 o;
            o.x = *reinterpret_cast<uint32_t*>(&p0);
            o.y = *reinterpret_cast<uint32_t*>(&p1);
            o.z = *reinterpret_cast<uint32_t*>(&p2);
            o.w = *reinterpret_cast<uint32_t*>(&p3);
            const size_t off = shadow_offset(dst_row0 + r, g * 8, d, rows_per_tile);
            *reinterpret_cast<uint4*>(shadow + off) = o;
        }
        ss = warp_butterfly_sum(ss);
        es = warp_butterfly_sum(es);
        if (center && !row_cdot) xs = warp_butterfly_sum(xs); else xs = ss;
        if (row_cdot) cd = warp_butterfly_sum(cd);
        if (lane == 0) {
            if (row_norms) row_norms[r] = sqrtf(ss);
            if (row_errs) row_errs[r] = (es == es) ? sqrtf(es) : INFINITY;
            if (row_cdot) row_cdot[r] = center ? cd : 0.0f;
            // non-negative floats order like their bit patterns; NaN/inf norms
            // saturate the bound (certificate then always fails -> exact path).
            if (maxnorm2_bits) {
                unsigned bits = (ss == ss) ? __float_as_uint(ss) : 0x7f800000u;
                atomicMax(maxnorm2_bits, bits);
                unsigned ebits = (es == es) ? __float_as_uint(es) : 0x7f800000u;
                atomicMax(maxnorm2_bits + 1, ebits);
                unsigned xbits = (xs == xs) ? __float_as_uint(xs) : 0x7f800000u;
                atomicMax(maxnorm2_bits + 2, xbits);
            }
        }
    }
}

// Column sums of rows [0, n_rows) (+ the sum of squared row norms in sums[d]): the centring decision of an index
// (api.cu::decide_center).  One block per slab of rows, one thread per column group; fp32 atomics into sums.
__global__ void __launch_bounds__(256)
column_sums_kernel(const float* __restrict__ src, int64_t n_rows, int d, int rows_per_block, float* __restrict__ sums) {
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    int64_t r1 = r0 + rows_per_block;
    if (r1 > n_rows) r1 = n_rows;
    float sq = 0.0f;
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
        float acc = 0.0f;
        int64_t r = r0;
        for (; r + 4 <= r1; r += 4) {  // four independent loads in flight
            const float v0 = src[r * (int64_t)d + c], v1 = src[(r + 1) * (int64_t)d + c];
            const float v2 = src[(r + 2) * (int64_t)d + c], v3 = src[(r + 3) * (int64_t)d + c];
            acc += (v0 + v1) + (v2 + v3);
            sq = fmaf(v0, v0, fmaf(v1, v1, fmaf(v2, v2, fmaf(v3, v3, sq))));
        }
        for (; r < r1; ++r) {
            const float v = src[r * (int64_t)d + c];
            acc += v;
            sq = fmaf(v, v, sq);
        }
        atomicAdd(sums + c, acc);
    }
    sq = warp_butterfly_sum(sq);
    if ((threadIdx.x & 31) == 0) atomicAdd(sums + d, sq);
}

// Row norms only (d not a multiple of 64: no shadow, exact path only).
__global__ void __launch_bounds__(256)
row_norms_kernel(const float* __restrict__ src, int64_t n_rows, int d,
                 unsigned* __restrict__ maxnorm2_bits, float* __restrict__ row_norms) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < n_rows; r += n_warps) {
        const float* row = src + r * (int64_t)d;
        float ss = 0.0f;
        for (int c = lane; c < d; c += 32) ss = fmaf(row[c], row[c], ss);
        ss = warp_butterfly_sum(ss);
        if (lane == 0) {
            if (row_norms) row_norms[r] = sqrtf(ss);
            if (maxnorm2_bits) {
                unsigned bits = (ss == ss) ? __float_as_uint(ss) : 0x7f800000u;
                atomicMax(maxnorm2_bits, bits);
            }
        }
    }
}

int launch_column_sums(const float* src, int64_t n_rows, int d, float* sums, cudaStream_t st) {
    if (n_rows <= 0) return 0;
    const int rows_per_block = 16;  // 512 blocks for the 8192-row sample of the centre decision
    column_sums_kernel<<<(unsigned)((n_rows + rows_per_block - 1) / rows_per_block), 256, 0, st>>>(src, n_rows, d,
                                                                                                    rows_per_block, sums);
    KIRAG_LAUNCH_OK("column_sums_kernel");
    return 0;
}

int launch_convert_rows(const float* src, int64_t n_rows, int d, int64_t dst_row0, void* shadow,
                        int rows_per_tile, unsigned* maxnorm2_bits, float* row_norms, float* row_errs,
                        const float* center, float* row_cdot, cudaStream_t st) {
    if (n_rows <= 0) return 0;
    const int threads = 256;
    const int64_t warps_needed = n_rows;
    int64_t blocks = (warps_needed * 32 + threads - 1) / threads;
    if (blocks > 148 * 16) blocks = 148 * 16;  // grid-stride, 16 CTAs per SM
    if (shadow) {
        KIRAG_CHECK(d % 64 == 0, "convert: d=%d is not a multiple of 64", d);
        KIRAG_CUDA_OK(launch_chained(convert_rows_kernel, dim3((unsigned)blocks), dim3(threads), 0, st, src, n_rows, d,
                                     dst_row0, (uint8_t*)shadow, rows_per_tile, maxnorm2_bits, row_norms, row_errs, center,
                                     row_cdot));
        KIRAG_LAUNCH_OK("convert_rows_kernel");
    } else {
        row_norms_kernel<<<(unsigned)blocks, threads, 0, st>>>(src, n_rows, d, maxnorm2_bits,
                                                              row_norms);
        KIRAG_LAUNCH_OK("row_norms_kernel");
    }
    return 0;
}

}  // namespace kirag
