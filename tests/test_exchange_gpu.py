"""The fused peer-memory exchange + merge kernel (csrc/exchange.cu) against the oracle's merge.

On one GPU the G "ranks" live in one process: G exchange buffers on the same device, connected by
raw pointers (kirag_exchange_connect_ptrs), one stream per rank so that the G kernels run
concurrently (each waits for the others' stores).  With >= 2 GPUs the real thing — one process
per GPU, CUDA IPC mappings, NVLink stores — runs through tools/exchange_check.py under torchrun.
"""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAD_D = np.float32(-3.4028234663852886e38)


def shard_lists(rng, G, nq, k, pad_last=0, n_scores=7):
    """Per-shard sorted result lists with many exact score ties and disjoint id ranges."""
    D = np.empty((G, nq, k), dtype=np.float32)
    I = np.empty((G, nq, k), dtype=np.int64)
    for g in range(G):
        for q in range(nq):
            ids = rng.choice(10**6, k, replace=False).astype(np.int64) + g * 10**6
            sc = rng.integers(-n_scores, n_scores + 1, size=k).astype(np.float32) * 0.5
            order = np.lexsort((ids, -sc))
            D[g, q], I[g, q] = sc[order], ids[order]
    if pad_last:
        D[-1, :, k - pad_last:] = PAD_D
        I[-1, :, k - pad_last:] = -1
    return D, I


class LocalRanks:
    """G exchange objects on device 0, wired to each other by raw pointers."""

    def __init__(self, G, max_nq, max_k):
        from kirag_b200 import _lib

        self.lib = _lib.load()
        self._lib_mod = _lib
        self.G = G
        self.h = []
        for r in range(G):
            h = ctypes.c_void_p()
            _lib.check(self.lib.kirag_exchange_create(0, r, G, max_nq, max_k, ctypes.byref(h)), "create")
            self.h.append(h)
        bufs = (ctypes.c_void_p * G)(*[self.lib.kirag_exchange_buffer(h) for h in self.h])
        for h in self.h:
            _lib.check(self.lib.kirag_exchange_connect_ptrs(h, bufs), "connect_ptrs")
        self.streams = [torch.cuda.Stream(device=0) for _ in range(G)]

    def merge(self, D_all, I_all):
        G, nq, k = D_all.shape
        Dd = [torch.from_numpy(D_all[g]).cuda() for g in range(G)]
        Id = [torch.from_numpy(I_all[g]).cuda() for g in range(G)]
        Do = [torch.empty((nq, k), dtype=torch.float32, device="cuda") for _ in range(G)]
        Io = [torch.empty((nq, k), dtype=torch.int64, device="cuda") for _ in range(G)]
        torch.cuda.synchronize()
        for g in range(G):
            self._lib_mod.check(
                self.lib.kirag_exchange_merge_topk(self.h[g], ctypes.c_void_p(Dd[g].data_ptr()),
                                                   ctypes.c_void_p(Id[g].data_ptr()), nq, k,
                                                   ctypes.c_void_p(Do[g].data_ptr()), ctypes.c_void_p(Io[g].data_ptr()),
                                                   ctypes.c_void_p(self.streams[g].cuda_stream)), "merge")
        torch.cuda.synchronize()
        return [d.cpu().numpy() for d in Do], [i.cpu().numpy() for i in Io]

    def merge_with_flags(self, D_all, I_all, flags_all):
        """flags_all: int32 [G, nq]; returns (ids per rank, any_flag per rank)."""
        G, nq, k = D_all.shape
        Dd = [torch.from_numpy(D_all[g]).cuda() for g in range(G)]
        Id = [torch.from_numpy(I_all[g]).cuda() for g in range(G)]
        Fd = [torch.from_numpy(flags_all[g]).cuda() for g in range(G)]
        Do = [torch.empty((nq, k), dtype=torch.float32, device="cuda") for _ in range(G)]
        Io = [torch.empty((nq, k), dtype=torch.int64, device="cuda") for _ in range(G)]
        torch.cuda.synchronize()
        for g in range(G):
            self._lib_mod.check(
                self.lib.kirag_exchange_merge_topk_flags(self.h[g], ctypes.c_void_p(Dd[g].data_ptr()),
                                                         ctypes.c_void_p(Id[g].data_ptr()), ctypes.c_void_p(Fd[g].data_ptr()),
                                                         nq, k, ctypes.c_void_p(Do[g].data_ptr()),
                                                         ctypes.c_void_p(Io[g].data_ptr()),
                                                         ctypes.c_void_p(self.streams[g].cuda_stream)), "merge_flags")
        torch.cuda.synchronize()
        return [i.cpu().numpy() for i in Io], [int(self.lib.kirag_exchange_last_any_flag(h)) for h in self.h]

    def close(self):
        for h in self.h:
            self.lib.kirag_exchange_destroy(h)


@pytest.mark.parametrize("G,nq,k,pad", [(2, 5, 10, 0), (4, 16, 100, 30), (8, 33, 100, 100), (3, 1, 1, 0),
                                        (8, 3, 1024, 300), (1, 7, 20, 5), (2, 64, 128, 0)])
def test_exchange_merge_matches_oracle(G, nq, k, pad):
    rng = np.random.default_rng(100 * G + nq + k)
    ranks = LocalRanks(G, max_nq=64, max_k=1024 if k > 128 else 128)
    try:
        # several calls on the same buffers: epochs, both parities, changing nq (stale flags of
        # blocks that a smaller call does not use)
        for call, nq_c in enumerate((nq, max(1, nq // 2), nq, nq)):
            D_all, I_all = shard_lists(rng, G, nq_c, k, pad_last=pad if call != 2 else 0)
            Do, Io = oracle.merge_topk(D_all, I_all)
            Dm, Im = ranks.merge(D_all, I_all)
            for g in range(G):  # every rank ends up with the same global answer
                assert np.array_equal(Im[g], Io), f"call {call} rank {g}: ids"
                assert np.array_equal(Dm[g], Do), f"call {call} rank {g}: scores"
    finally:
        ranks.close()


def test_exchange_short_union_pads_like_faiss():
    """Fewer than k results in the whole union: tail is (-FLT_MAX, -1)."""
    rng = np.random.default_rng(5)
    G, nq, k = 3, 4, 12
    D_all, I_all = shard_lists(rng, G, nq, k)
    D_all[:, :, 3:] = PAD_D
    I_all[:, :, 3:] = -1
    ranks = LocalRanks(G, 8, 16)
    try:
        Dm, Im = ranks.merge(D_all, I_all)
    finally:
        ranks.close()
    Do, Io = oracle.merge_topk(D_all, I_all)
    assert np.array_equal(Im[0], Io) and np.array_equal(Dm[0], Do)
    assert np.all(Im[0][:, 9:] == -1) and np.all(Dm[0][:, 9:] == PAD_D)


def test_exchange_carries_certificate_flags_to_every_rank():
    """kirag_exchange_merge_topk_flags: one rank's flagged query makes the OR word non-zero on EVERY rank (all ranks
    then re-run the exchange together); all-zero flags leave it zero; the merged rows are unaffected."""
    rng = np.random.default_rng(11)
    G, nq, k = 4, 37, 20
    ranks = LocalRanks(G, max_nq=64, max_k=128)
    try:
        for call, (fr, fq) in enumerate(((None, None), (2, 36), (None, None), (0, 0), (3, 5))):
            D_all, I_all = shard_lists(rng, G, nq, k)
            flags = np.zeros((G, nq), dtype=np.int32)
            if fr is not None:
                flags[fr, fq] = 1 + call % 2
            Io = oracle.merge_topk(D_all, I_all)[1]
            Im, anyf = ranks.merge_with_flags(D_all, I_all, flags)
            assert all(np.array_equal(Im[g], Io) for g in range(G)), f"call {call}"
            assert anyf == [0 if fr is None else 1] * G, f"call {call}: {anyf}"
            # a flag-less exchange in between keeps working on the same buffers (parity alternates)
            Dm2, Im2 = ranks.merge(D_all, I_all)
            assert all(np.array_equal(Im2[g], Io) for g in range(G))
    finally:
        ranks.close()


def test_exchange_orders_calls_across_streams_and_replays_from_a_graph():
    """Epoch double-buffering is only safe if a rank's exchanges execute in call order: a call on another stream
    waits for the previous exchange (event), and because the epoch lives on the device a captured exchange can be
    replayed from a CUDA graph any number of times."""
    from kirag_b200 import _lib

    lib = _lib.load()
    h = ctypes.c_void_p()
    _lib.check(lib.kirag_exchange_create(0, 0, 1, 8, 8, ctypes.byref(h)), "create")
    try:
        D = torch.zeros((2, 4), device="cuda")
        I = torch.arange(8, device="cuda").reshape(2, 4)
        Do, Io = torch.empty_like(D), torch.empty_like(I)
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        args = (ctypes.c_void_p(D.data_ptr()), ctypes.c_void_p(I.data_ptr()), 2, 4, ctypes.c_void_p(Do.data_ptr()),
                ctypes.c_void_p(Io.data_ptr()))
        for st in (s1, s2, s1, s2, s2, s1):
            _lib.check(lib.kirag_exchange_merge_topk(h, *args, ctypes.c_void_p(st.cuda_stream)), "merge")
        torch.cuda.synchronize()
        assert torch.equal(Io, I)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(s1):
            with torch.cuda.graph(graph, stream=s1):
                _lib.check(lib.kirag_exchange_merge_topk(h, *args, ctypes.c_void_p(s1.cuda_stream)), "merge (captured)")
            for trial in range(5):  # both parities, several times
                I.add_(100)
                Io.zero_()
                graph.replay()
                s1.synchronize()
                assert torch.equal(Io, I), trial
                assert lib.kirag_exchange_last_any_flag(h) == 0
        _lib.check(lib.kirag_exchange_merge_topk(h, *args, ctypes.c_void_p(s2.cuda_stream)), "merge after replays")
        torch.cuda.synchronize()
        assert torch.equal(Io, I)
    finally:
        lib.kirag_exchange_destroy(h)


def test_exchange_rejects_bad_arguments():
    from kirag_b200 import _lib

    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.kirag_exchange_create(0, 2, 2, 8, 8, ctypes.byref(h)) != 0  # rank out of range
    assert lib.kirag_exchange_create(0, 0, 17, 8, 8, ctypes.byref(h)) != 0  # too many ranks
    _lib.check(lib.kirag_exchange_create(0, 0, 2, 8, 8, ctypes.byref(h)), "create")
    try:
        t = torch.zeros(64, device="cuda")
        p = ctypes.c_void_p(t.data_ptr())
        # peers not connected yet
        assert lib.kirag_exchange_merge_topk(h, p, p, 2, 4, p, p, None) != 0
        assert "not connected" in _lib.last_error()
    finally:
        lib.kirag_exchange_destroy(h)


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_process_peer_exchange_equals_nccl_and_unsharded():
    """One process per GPU, CUDA IPC + NVLink stores; checked inside tools/exchange_check.py."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29653", os.path.join(ROOT, "tools", "exchange_check.py")]
    proc = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    assert "exchange_check ok" in proc.stdout
