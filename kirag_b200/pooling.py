"""Embedding epilogue: masked mean-pool / CLS-select + L2-normalise in one kernel.

Drop-ins for
    average_pool(last_hidden_states, attention_mask)   /root/reference/retriever/encoders.py:56-58
                                                       /root/reference/retriever/e5.py:46-48
    E5Encoder.forward tail  (average_pool + F.normalize)   encoders.py:75-76,  e5.py:59-60
    BGEEncoder.forward tail (hidden[:, 0] + F.normalize)   encoders.py:115-117
    ContrieverEncoder tail  (average_pool only)            encoders.py:94-95

PyTorch is used for tensor memory, streams and autograd bookkeeping only; the
arithmetic is kirag_pool_normalize* in the sm_100a library.  CPU tensors are
rejected — there is no fallback.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

_HIDDEN_DTYPES = {torch.float32: _lib.DTYPE_F32, torch.bfloat16: _lib.DTYPE_BF16, torch.float16: _lib.DTYPE_F16}
_MASK_DTYPES = {torch.int64: _lib.MASK_I64, torch.int32: _lib.MASK_I32, torch.uint8: _lib.MASK_U8,
                torch.bool: _lib.MASK_U8}


def _p(t) -> ctypes.c_void_p:
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _prep(hidden: torch.Tensor, mask, mode: int):
    if not hidden.is_cuda:
        raise RuntimeError("kirag_b200.pooling: hidden states must be a CUDA tensor (no CPU fallback)")
    if hidden.dim() != 3:
        raise AssertionError("hidden states must be [B, S, H]")
    if hidden.dtype not in _HIDDEN_DTYPES:
        raise TypeError(f"unsupported hidden dtype {hidden.dtype}")
    if hidden.stride(2) != 1:
        hidden = hidden.contiguous()
    if mode == _lib.POOL_MEAN:
        if mask is None:
            raise AssertionError("mean pooling needs an attention mask")
        if mask.dtype not in _MASK_DTYPES:
            mask = mask.to(torch.int64)
        if mask.device != hidden.device:
            mask = mask.to(hidden.device)
        if mask.dim() != 2 or mask.shape[0] != hidden.shape[0] or mask.shape[1] != hidden.shape[1]:
            raise AssertionError("attention mask must be [B, S]")
        if mask.stride(1) != 1:
            mask = mask.contiguous()
    return hidden, mask


class _PoolNormalize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hidden, mask, mode: int, normalize: bool, typed: bool):
        hidden, mask = _prep(hidden, mask, mode)
        B, S, H = hidden.shape
        lib = _lib.load()
        out = torch.empty((B, H), dtype=torch.float32, device=hidden.device)
        norms = torch.empty((B,), dtype=torch.float32, device=hidden.device)
        # bf16 / fp16 hidden states (autocast): the kernel also writes the result in that dtype — no cast kernel
        typed = bool(typed) and hidden.dtype != torch.float32
        out_typed = torch.empty((B, H), dtype=hidden.dtype, device=hidden.device) if typed else None
        st = torch.cuda.current_stream(hidden.device).cuda_stream
        mdt = _MASK_DTYPES[mask.dtype] if mask is not None else _lib.MASK_I64
        mb = mask.stride(0) if mask is not None else 0
        if B > 0:
            if typed:
                _lib.check(
                    lib.kirag_pool_normalize_typed(
                        _p(hidden), _p(mask), _p(out), _p(out_typed), _p(norms), B, S, H, hidden.stride(0),
                        hidden.stride(1), mb, _HIDDEN_DTYPES[hidden.dtype], mdt, mode, int(bool(normalize)),
                        hidden.device.index, ctypes.c_void_p(st)),
                    "pool_normalize_typed")
            else:
                _lib.check(
                    lib.kirag_pool_normalize_fwd_saved(
                        _p(hidden), _p(mask), _p(out), _p(norms), B, S, H, hidden.stride(0), hidden.stride(1), mb,
                        _HIDDEN_DTYPES[hidden.dtype], mdt, mode, int(bool(normalize)), hidden.device.index,
                        ctypes.c_void_p(st)),
                    "pool_normalize")
        ctx.mode, ctx.normalize = mode, bool(normalize)
        ctx.hidden_dtype, ctx.shape = hidden.dtype, (B, S, H)
        ctx.save_for_backward(out, norms, mask if mask is not None else torch.empty(0, device=hidden.device))
        ctx.has_mask = mask is not None
        return out_typed if typed else out

    @staticmethod
    def backward(ctx, grad_out):
        out, norms, mask = ctx.saved_tensors
        B, S, H = ctx.shape
        lib = _lib.load()
        grad_out = grad_out.contiguous().float()
        grad_hidden = torch.empty((B, S, H), dtype=ctx.hidden_dtype, device=grad_out.device)
        st = torch.cuda.current_stream(grad_out.device).cuda_stream
        m = mask if ctx.has_mask else None
        mdt = _MASK_DTYPES[m.dtype] if m is not None else _lib.MASK_I64
        mb = m.stride(0) if m is not None else 0
        if B > 0:
            _lib.check(
                lib.kirag_pool_normalize_backward(
                    _p(grad_out), _p(out), _p(norms), _p(m), _p(grad_hidden), B, S, H, mb,
                    _HIDDEN_DTYPES[ctx.hidden_dtype], mdt, ctx.mode, int(ctx.normalize), grad_out.device.index,
                    ctypes.c_void_p(st)),
                "pool_normalize_backward")
        return grad_hidden, None, None, None, None


def pool_normalize(last_hidden_states: torch.Tensor, attention_mask=None, mode: str = "mean",
                   normalize: bool = True, out_dtype=None) -> torch.Tensor:
    """Fused epilogue.  mode 'mean' (E5) or 'cls' (BGE).  Returns [B, H]; dtype follows the
    hidden states (as the reference's ops do) unless out_dtype is given."""
    m = {"mean": _lib.POOL_MEAN, "cls": _lib.POOL_CLS}[mode]
    want = last_hidden_states.dtype if out_dtype is None else out_dtype
    out = _PoolNormalize.apply(last_hidden_states, attention_mask, m, normalize, want == last_hidden_states.dtype)
    return out if out.dtype == want else out.to(want)


def average_pool(last_hidden_states: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
    """Same signature and semantics as the reference's average_pool (encoders.py:56-58)."""
    return pool_normalize(last_hidden_states, attention_mask, mode="mean", normalize=False)


def e5_embed(last_hidden_states: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
    """F.normalize(average_pool(h, m), p=2, dim=1)  — E5Encoder.forward tail (encoders.py:75-76)."""
    return pool_normalize(last_hidden_states, attention_mask, mode="mean", normalize=True)


def bge_embed(last_hidden_states: torch.Tensor) -> torch.Tensor:
    """F.normalize(h[:, 0], p=2, dim=1) — BGEEncoder.forward tail (encoders.py:115-117)."""
    return pool_normalize(last_hidden_states, None, mode="cls", normalize=True)


def patch_reference_encoders(encoders_module=None, e5_module=None):
    """Substitute the fused epilogue into the reference's modules; returns a callable that undoes it.

    `encoders_module` is the imported /root/reference/retriever/encoders.py,
    `e5_module` the imported retriever/e5.py.  E5Encoder / BGEEncoder keep their
    HF BertModel body (out of scope) and get a forward whose tail is one kernel
    (encoders.py:67-77 and :106-118 otherwise unchanged: same arguments, same return).
    """
    from transformers import BertModel

    saved = []

    def swap(obj, name, value):
        saved.append((obj, name, getattr(obj, name)))
        setattr(obj, name, value)

    if encoders_module is not None:
        def e5_forward(self, input_ids, attention_mask, token_type_ids=None, **kwargs):
            out = BertModel.forward(self, input_ids=input_ids, attention_mask=attention_mask,
                                    token_type_ids=token_type_ids, return_dict=True)
            return e5_embed(out.last_hidden_state, attention_mask)

        def bge_forward(self, input_ids, attention_mask, token_type_ids=None, **kwargs):
            out = BertModel.forward(self, input_ids=input_ids, attention_mask=attention_mask,
                                    token_type_ids=token_type_ids, return_dict=True)
            return bge_embed(out.last_hidden_state)

        swap(encoders_module, "average_pool", average_pool)
        swap(encoders_module.E5Encoder, "forward", e5_forward)
        swap(encoders_module.BGEEncoder, "forward", bge_forward)
    if e5_module is not None:
        swap(e5_module, "average_pool", average_pool)

    def undo():
        while saved:
            obj, name, value = saved.pop()
            setattr(obj, name, value)

    return undo
