// rescore.cu — fp32 rescoring of the candidates that survived the bf16 filter.
//
// The bf16 tcgen05 scan only decides WHICH rows can be in the top-k; the
// score that is returned (and that orders the result) is recomputed here from
// the fp32 master with the canonical summation order, so it is what
// faiss.IndexFlatIP.search would return up to fp32 summation order (reference
// call site /root/reference/retriever/index.py:47).
//
// One warp per (query, candidate): a gather of 4*d-byte rows.  Implementation
// traffic (not algorithmic): nq * m * d * 4 bytes.
#include "common.cuh"

namespace kirag {

__global__ void __launch_bounds__(256)
rescore_kernel(const float* __restrict__ master, int d, const float* __restrict__ q,
               const Cand* __restrict__ cand, const int* __restrict__ cnt, int cand_stride, int m,
               float* __restrict__ out, int64_t nq, bool vec4, const float* __restrict__ tauk,
               const float* __restrict__ qnorm, const float* __restrict__ qerr, float eps_a, float eps_b) {
    pdl_wait();
    pdl_launch_dependents();
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= nq * m) return;
    const int64_t qi = warp / m;
    const int j = (int)(warp - qi * m);
    int count = cnt ? cnt[qi] : m;
    if (count > m) count = m;
    float result = __int_as_float(0x7fc00000);  // NaN marks an empty slot
    if (j < count) {
        const Cand cd = cand[qi * (int64_t)cand_stride + j];
        const int32_t row = cd.id;
        // a candidate this far below the k-th best approximate score cannot reach the top-k (select.cu, tauk)
        bool needed = true;
        if (tauk) needed = cd.s >= __ldcg(tauk + qi) - 2.0f * (eps_a * __ldcg(qnorm + qi) + eps_b * __ldcg(qerr + qi));
        if (row >= 0 && needed) {
            const float* x = master + (int64_t)row * d;
            const float part = canonical_partial(x, q + qi * (int64_t)d, d, lane, vec4);
            result = warp_butterfly_sum(part);
        }
    }
    if (lane == 0) out[qi * (int64_t)m + j] = result;
}

int launch_rescore(const float* master, int d, const float* q, const Cand* cand, const int* cnt,
                   int cand_stride, int m, float* out_scores, int64_t nq, const float* tauk, const float* qnorm,
                   const float* qerr, float eps_a, float eps_b, cudaStream_t st) {
    if (nq <= 0 || m <= 0) return 0;
    const int threads = 256;
    const int64_t warps = nq * m;
    const int64_t blocks = (warps + (threads / 32) - 1) / (threads / 32);
    KIRAG_CHECK(blocks < 0x7fffffffLL, "rescore: too many candidates (%lld warps)", (long long)warps);
    const bool vec4 = (d % 4 == 0) && ((reinterpret_cast<uintptr_t>(master) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(q) & 15) == 0);
    KIRAG_CUDA_OK(launch_chained(rescore_kernel, dim3((unsigned)blocks), dim3(threads), 0, st, master, d, q, cand, cnt,
                                 cand_stride, m, out_scores, nq, vec4, tauk, qnorm, qerr, eps_a, eps_b));
    KIRAG_LAUNCH_OK("rescore_kernel");
    return 0;
}

}  // namespace kirag
