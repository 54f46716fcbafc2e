"""Tensor-level entry points over the C ABI: build ``daam_layer`` descriptors from torch tensors and launch.

This is the thinnest layer above ``libdaam_b200.so``: no state, no policy. ``trace.py`` uses it from the attention
hook; the parity tests and ``bench.py`` call it directly with Q/K tensors.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from . import _native

__all__ = ['cond_half', 'make_layer_desc', 'new_accumulator', 'pack', 'accumulate', 'accumulate_layer',
           'attention_probs', 'accumulate_probs']

_DTYPES = {torch.float32: _native.DAAM_F32, torch.float16: _native.DAAM_F16, torch.bfloat16: _native.DAAM_BF16}


def cond_half(bsz: int, heads: int) -> Tuple[int, int, int, int]:
    """Which slice of the ``batch x heads`` axis the reference keeps (``map_[map_.size(0) // 2:]``, daam/trace.py:240).

    Returns ``(first_sample, n_prompts, first_head, n_heads)``: for a CFG batch ``[uncond x N, cond x N]`` the N
    conditional samples with all heads; for a lone sample (no guidance) the upper half of its heads -- the reference's
    behaviour, kept faithfully. Other odd batch sizes cut through a sample and are rejected.
    """
    if bsz % 2 == 0:
        return bsz // 2, bsz // 2, 0, heads
    if bsz == 1:
        return 0, 1, heads // 2, heads - heads // 2
    raise RuntimeError(f'a batch of {bsz} is neither a CFG pair batch nor a single sample')


def new_accumulator(n_prompts: int, heads: int, hw: int, device) -> torch.Tensor:
    return torch.zeros((n_prompts, heads, _native.TOKENS, hw), dtype=torch.float32, device=device)


def make_layer_desc(q: torch.Tensor, k: torch.Tensor, acc: torch.Tensor, heads: int, scale: float
                    ) -> _native.DaamLayer:
    """``q [B, hw, heads*d]`` / ``k [B, 77, heads*d]`` as ``to_q`` / ``to_k`` emit them (last axis contiguous) and the
    fp32 accumulator ``[n_prompts, n_heads, 77, hw]`` of the kept slice -> one ``daam_layer``."""
    if not (q.is_cuda and k.is_cuda and acc.is_cuda):
        raise RuntimeError('daam_b200 computes on CUDA tensors only (there is no CPU fallback)')
    if q.dtype not in _DTYPES or k.dtype != q.dtype:
        raise RuntimeError(f'unsupported projection dtypes {q.dtype}/{k.dtype}')
    if q.stride(-1) != 1 or k.stride(-1) != 1:
        raise RuntimeError('the channel axis of q and k must be contiguous')
    bsz, hw, chan = q.shape
    d = chan // heads
    first, n_prompts, head0, n_heads = cond_half(bsz, heads)
    if tuple(acc.shape) != (n_prompts, n_heads, _native.TOKENS, hw) or acc.dtype != torch.float32 \
            or not acc.is_contiguous():
        raise RuntimeError(f'accumulator must be contiguous fp32 {(n_prompts, n_heads, _native.TOKENS, hw)}, '
                           f'got {acc.dtype} {tuple(acc.shape)}')
    es = q.element_size()
    return _native.DaamLayer(
        q=q.data_ptr() + (first * q.stride(0) + head0 * d) * es,
        k=k.data_ptr() + (first * k.stride(0) + head0 * d) * es,
        acc=acc.data_ptr(),
        q_stride_prompt=q.stride(0), q_stride_pixel=q.stride(1), q_stride_head=d,
        k_stride_prompt=k.stride(0), k_stride_token=k.stride(1), k_stride_head=d,
        n_prompts=n_prompts, heads=n_heads, hw=hw, tokens=k.shape[1], head_dim=d,
        dtype=_DTYPES[q.dtype], scale=float(scale), reserved=0)


def pack(descs: Sequence[_native.DaamLayer]) -> _native.PackedLayers:
    """Pre-build the host-side ``daam_layer[]`` once for layer calls that are replayed (bench loops, CUDA graphs)."""
    return _native.PackedLayers(list(descs))


def accumulate(descs, device, stream: Optional[torch.cuda.Stream] = None, flags: int = _native.ACC_AUTO):
    """Enqueue the fused kernel over the given layer calls (sequence of descriptors or :func:`pack` result) on
    ``stream`` (default: the current stream of ``device``)."""
    dev = device if isinstance(device, torch.device) else torch.device(device)
    index = dev.index if dev.index is not None else torch.cuda.current_device()
    s = torch.cuda.current_stream(index) if stream is None else stream
    if index == torch.cuda.current_device():
        _native.accumulate(descs, s.cuda_stream, flags)
    else:
        with torch.cuda.device(index):
            _native.accumulate(descs, s.cuda_stream, flags)


def accumulate_layer(q: torch.Tensor, k: torch.Tensor, heads: int, scale: Optional[float] = None,
                     acc: Optional[torch.Tensor] = None, flags: int = _native.ACC_AUTO) -> torch.Tensor:
    """One layer call on the current stream; allocates the accumulator when none is given. Returns it."""
    bsz, hw, chan = q.shape
    _, n_prompts, _, n_heads = cond_half(bsz, heads)
    if acc is None:
        acc = new_accumulator(n_prompts, n_heads, hw, q.device)
    if scale is None:
        scale = (chan // heads) ** -0.5
    accumulate([make_layer_desc(q, k, acc, heads, scale)], q.device, flags=flags)
    return acc


def attention_probs(q: torch.Tensor, k: torch.Tensor, heads: int, scale: Optional[float] = None) -> torch.Tensor:
    """Materialised ``softmax(scale * Q K^T)`` for EVERY sample: ``[B*heads, hw, 77]`` in the dtype of ``q`` -- what
    diffusers' ``get_attention_scores`` returns at daam/trace.py:276 (rows ordered ``b*heads + head``). Compatibility
    path for ``save_heads``; the traced hot path never materialises this tensor."""
    if not (q.is_cuda and k.is_cuda):
        raise RuntimeError('daam_b200 computes on CUDA tensors only (there is no CPU fallback)')
    if q.dtype not in _DTYPES or k.dtype != q.dtype:
        raise RuntimeError(f'unsupported projection dtypes {q.dtype}/{k.dtype}')
    if q.stride(-1) != 1 or k.stride(-1) != 1:
        q, k = q.contiguous(), k.contiguous()
    bsz, hw, chan = q.shape
    d = chan // heads
    if scale is None:
        scale = d ** -0.5
    probs = torch.empty((bsz * heads, hw, k.shape[1]), dtype=q.dtype, device=q.device)
    desc = _native.DaamLayer(
        q=q.data_ptr(), k=k.data_ptr(), acc=None,
        q_stride_prompt=q.stride(0), q_stride_pixel=q.stride(1), q_stride_head=d,
        k_stride_prompt=k.stride(0), k_stride_token=k.stride(1), k_stride_head=d,
        n_prompts=bsz, heads=heads, hw=hw, tokens=k.shape[1], head_dim=d,
        dtype=_DTYPES[q.dtype], scale=float(scale), reserved=0)
    with torch.cuda.device(q.device):
        _native.attention_probs(desc, probs.data_ptr(), torch.cuda.current_stream(q.device).cuda_stream)
    return probs


def accumulate_probs(probs: torch.Tensor, acc: torch.Tensor):
    """``acc[r][t][pixel] += probs[first + r][pixel][t]`` with ``first = rows // 2``: the reference's "second half of
    the batch*heads axis" (daam/trace.py:240) applied to supplied probabilities (``load_heads``, trace.py:281-294).
    ``acc``: fp32 ``[n_prompts, n_heads, 77, hw]`` (its leading two axes flatten to the kept rows)."""
    if not (probs.is_cuda and acc.is_cuda):
        raise RuntimeError('daam_b200 computes on CUDA tensors only (there is no CPU fallback)')
    probs = probs.contiguous()
    rows, hw, tokens = probs.shape
    first = rows // 2
    kept = rows - first
    if acc.dtype != torch.float32 or not acc.is_contiguous() or acc.numel() != kept * tokens * hw:
        raise RuntimeError(f'accumulator must be contiguous fp32 with {kept} x {tokens} x {hw} elements')
    with torch.cuda.device(probs.device):
        _native.accumulate_probs(probs.data_ptr(), _DTYPES[probs.dtype], first, kept, hw, tokens, acc.data_ptr(),
                                 torch.cuda.current_stream(probs.device).cuda_stream)
