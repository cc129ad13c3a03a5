// scan_exact.cu — exact fp32 scan of the master on CUDA cores.
//
// The arithmetic of faiss.IndexFlatIP.search (reference call site
// /root/reference/retriever/index.py:47): every query against every corpus
// row, fp32 multiply-accumulate.  This is the certified-exact path: it is used
// for shapes the tcgen05 filter cannot take (d % 64 != 0), when the caller
// forces KIRAG_PATH_EXACT, and for the rare query whose filter-path result
// fails the exactness certificate.  Scores use the canonical summation order
// (common.cuh) so they are bit-identical to the rescoring kernel's.
//
// HBM-bound: each 4*d-byte row is read once for a group of kExactNQ (8) queries.
// Algorithmic bytes per launch = n*d*4 (+ kExactNQ*n*4 of scores written).
#include "common.cuh"

namespace kirag {

template <bool VEC4>
__global__ void __launch_bounds__(256)
scan_exact_kernel(const float* __restrict__ master, int64_t n, int d,
                  const float* __restrict__ q, int nq_valid, float* __restrict__ scores,
                  int64_t ld) {
    extern __shared__ float sq[];  // [kExactNQ][d]
    for (int i = threadIdx.x; i < kExactNQ * d; i += blockDim.x) {
        const int qi = i / d;
        sq[i] = (qi < nq_valid) ? q[i] : 0.0f;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    // each warp owns groups of 32 consecutive rows so the score store is coalesced
    for (int64_t base = warp * 32; base < n; base += n_warps * 32) {
        float keep[kExactNQ];
#pragma unroll
        for (int qi = 0; qi < kExactNQ; ++qi) keep[qi] = 0.0f;
        const int rows = (int)((n - base < 32) ? (n - base) : 32);
        // two rows per trip: every query float4 read from shared memory feeds both rows.  With one row per trip and 8
        // queries the kernel was bound by shared-memory bandwidth (64 LDS.128 per row and lane: 256 cycles per row
        // against the 175 cycles per row that the HBM rate allows per SM)
        for (int r = 0; r < rows; r += 2) {
            const int rb = (r + 1 < rows) ? r + 1 : r;  // odd tail: the second row repeats the first, its result is unused
            const float* xa = master + (base + r) * (int64_t)d;
            const float* xb = master + (base + rb) * (int64_t)d;
            float acc[2][kExactNQ];
#pragma unroll
            for (int qi = 0; qi < kExactNQ; ++qi) acc[0][qi] = acc[1][qi] = 0.0f;
            if (VEC4) {
                // the loads of both rows are issued in batches of 4 + 4 before the FMAs that consume them; the FMA
                // chain of every (row, query) pair — and with it the canonical summation order — is unchanged
                constexpr int U = 4;
                int c = lane * 4;
                for (; c + (U - 1) * 128 < d; c += U * 128) {
                    float4 va[U], vb[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        va[u] = __ldcs(reinterpret_cast<const float4*>(xa + c + u * 128));
                        vb[u] = __ldcs(reinterpret_cast<const float4*>(xb + c + u * 128));
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
#pragma unroll
                        for (int qi = 0; qi < kExactNQ; ++qi) {
                            const float4 qv = *reinterpret_cast<const float4*>(sq + qi * d + c + u * 128);
                            acc[0][qi] = __fmaf_rn(va[u].x, qv.x, acc[0][qi]);
                            acc[0][qi] = __fmaf_rn(va[u].y, qv.y, acc[0][qi]);
                            acc[0][qi] = __fmaf_rn(va[u].z, qv.z, acc[0][qi]);
                            acc[0][qi] = __fmaf_rn(va[u].w, qv.w, acc[0][qi]);
                            acc[1][qi] = __fmaf_rn(vb[u].x, qv.x, acc[1][qi]);
                            acc[1][qi] = __fmaf_rn(vb[u].y, qv.y, acc[1][qi]);
                            acc[1][qi] = __fmaf_rn(vb[u].z, qv.z, acc[1][qi]);
                            acc[1][qi] = __fmaf_rn(vb[u].w, qv.w, acc[1][qi]);
                        }
                    }
                }
                for (; c < d; c += 128) {
                    const float4 va = __ldcs(reinterpret_cast<const float4*>(xa + c));
                    const float4 vb = __ldcs(reinterpret_cast<const float4*>(xb + c));
#pragma unroll
                    for (int qi = 0; qi < kExactNQ; ++qi) {
                        const float4 qv = *reinterpret_cast<const float4*>(sq + qi * d + c);
                        acc[0][qi] = __fmaf_rn(va.x, qv.x, acc[0][qi]);
                        acc[0][qi] = __fmaf_rn(va.y, qv.y, acc[0][qi]);
                        acc[0][qi] = __fmaf_rn(va.z, qv.z, acc[0][qi]);
                        acc[0][qi] = __fmaf_rn(va.w, qv.w, acc[0][qi]);
                        acc[1][qi] = __fmaf_rn(vb.x, qv.x, acc[1][qi]);
                        acc[1][qi] = __fmaf_rn(vb.y, qv.y, acc[1][qi]);
                        acc[1][qi] = __fmaf_rn(vb.z, qv.z, acc[1][qi]);
                        acc[1][qi] = __fmaf_rn(vb.w, qv.w, acc[1][qi]);
                    }
                }
            } else {
                for (int c = lane * 4; c < d; c += 128) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (c + j < d) {
                            const float va = xa[c + j], vb = xb[c + j];
#pragma unroll
                            for (int qi = 0; qi < kExactNQ; ++qi) {
                                const float qv = sq[qi * d + c + j];
                                acc[0][qi] = __fmaf_rn(va, qv, acc[0][qi]);
                                acc[1][qi] = __fmaf_rn(vb, qv, acc[1][qi]);
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int qi = 0; qi < kExactNQ; ++qi) {
                const float ta = warp_butterfly_sum(acc[0][qi]);
                const float tb = warp_butterfly_sum(acc[1][qi]);
                if (lane == r) keep[qi] = ta;
                if (rb != r && lane == rb) keep[qi] = tb;
            }
        }
        if (lane < rows) {
#pragma unroll
            for (int qi = 0; qi < kExactNQ; ++qi)
                if (qi < nq_valid) scores[qi * ld + base + lane] = keep[qi];
        }
    }
}

int launch_scan_exact(const float* master, int64_t n, int d, const float* q, int nq_valid,
                      float* scores, int64_t ld, int num_sms, cudaStream_t st) {
    if (n <= 0) return 0;
    const int threads = 256;
    const size_t smem = (size_t)kExactNQ * d * sizeof(float);
    KIRAG_CHECK(smem <= 200 * 1024, "scan_exact: d=%d too large for the query tile", d);
    int64_t groups = (n + 31) / 32;
    int64_t blocks = (groups + (threads / 32) - 1) / (threads / 32);
    const int64_t max_blocks = (int64_t)num_sms * 6;
    if (blocks > max_blocks) blocks = max_blocks;
    const bool vec4 = (d % 4 == 0) && ((reinterpret_cast<uintptr_t>(master) & 15) == 0);
    if (vec4) {
        if (smem > 48 * 1024 && ensure_dynamic_smem(scan_exact_kernel<true>, smem)) return 1;
        scan_exact_kernel<true><<<(unsigned)blocks, threads, smem, st>>>(master, n, d, q, nq_valid,
                                                                         scores, ld);
    } else {
        if (smem > 48 * 1024 && ensure_dynamic_smem(scan_exact_kernel<false>, smem)) return 1;
        scan_exact_kernel<false><<<(unsigned)blocks, threads, smem, st>>>(master, n, d, q, nq_valid,
                                                                          scores, ld);
    }
    KIRAG_LAUNCH_OK("scan_exact_kernel");
    return 0;
}

}  // namespace kirag
