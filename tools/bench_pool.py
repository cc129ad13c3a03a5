"""Pooling epilogue benchmark (BASELINE configs[4]: e5-large-v2 hidden states, seq 512).

Times kirag_pool_normalize (one kernel) against the reference's own ATen expression
(retriever/encoders.py:56-58,76 run on the same GPU) and reports achieved HBM GB/s over the
ALGORITHMIC bytes: sum_b len_b*H*sizeof(hidden) + B*S*8 (int64 mask) + B*H*4 (output).
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from kirag_b200 import pooling  # noqa: E402
from oracle import oracle  # noqa: E402  (the reference expression, as the timed ATen baseline and the checker)


def timeit(fn, iters=20, warmup=5, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(iters):
        if flush is not None:
            flush.sum()  # evict L2 between iterations by READING a buffer larger than the 126 MB L2 (leaves clean lines)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    times.sort()
    return times[len(times) // 2], times[0]


def main():
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    dev = torch.device("cuda", 0)
    flush = torch.zeros(64 << 20, dtype=torch.int32, device=dev)
    out = []
    only = os.environ.get("POOL_ONLY")
    for idx, (B, S, H, dtype, ragged) in enumerate([(256, 512, 1024, torch.float32, False), (256, 512, 1024, torch.float32, True),
                                     (256, 512, 1024, torch.bfloat16, False), (8, 512, 1024, torch.float32, False),
                                     (4, 128, 1024, torch.float32, True), (1024, 512, 1024, torch.bfloat16, False)]):
        if only is not None and int(only) != idx:
            continue
        g = torch.Generator(device=dev)
        g.manual_seed(777)
        h = torch.randn(B, S, H, generator=g, device=dev, dtype=torch.float32).to(dtype)
        if ragged:
            gl = torch.Generator().manual_seed(778)
            lens = torch.randint(1, S + 1, (B,), generator=gl)
        else:
            lens = torch.full((B,), S)
        m = (torch.arange(S)[None, :] < lens[:, None]).to(torch.int64).to(dev)
        alg = int(lens.sum()) * H * h.element_size() + B * S * 8 + B * H * 4
        ours = lambda: pooling.pool_normalize(h, m, out_dtype=torch.float32)
        ref = lambda: oracle.pool_normalize_torch(h, m)
        err = float((ours().float() - ref().float()).abs().max())
        t_ours, t_ours_min = timeit(ours, flush=flush)
        t_ref, _ = timeit(ref, flush=flush)
        gbs = alg / (t_ours * 1e-3) / 1e9
        rec = {"B": B, "S": S, "H": H, "dtype": str(dtype).split(".")[-1], "ragged": ragged,
               "algorithmic_bytes": alg, "ours_ms": t_ours, "ours_ms_min": t_ours_min, "aten_reference_ms": t_ref,
               "speedup_vs_aten": t_ref / t_ours, "achieved_gbs": gbs, "frac_of_measured_hbm": gbs / peaks["hbm_gbs"],
               "max_abs_err_vs_reference": err}
        out.append(rec)
        print(json.dumps(rec), flush=True)
    return out


if __name__ == "__main__":
    main()
