"""Mask overlap measures of the reference's evaluation helpers (``/root/reference/daam/evaluate.py:14-35``).

Only ``compute_iou`` / ``compute_ioa`` are mirrored -- the part of SURVEY.md section 8f rank 4 that touches heat maps.
When the two masks differ in size the reference bicubically resizes the first to the second and binarises it at 1;
that resize runs on the native ``daam_expand_as`` kernel (absolute mode). The COCO evaluators (``UnsupervisedEvaluator``,
``MeanEvaluator``, PNG mask loading) are out of scope.
"""
from __future__ import annotations

import torch

from . import _native

__all__ = ['compute_iou', 'compute_ioa']


def _match_size(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """evaluate.py:15-18 / 27-30: bicubic resize of ``a`` to ``b``'s shape, then ``a < 1 -> 0``, ``a >= 1 -> 1``."""
    if a.shape[0] == b.shape[0]:
        return a
    if not a.is_cuda:
        raise RuntimeError('compute_iou/compute_ioa resize on CUDA tensors only (there is no CPU fallback)')
    if a.shape[0] != a.shape[1]:
        raise ValueError('the native resize takes square source maps')
    src = a.detach().float().contiguous()
    out = torch.empty(tuple(b.shape), dtype=torch.float32, device=a.device)
    scratch = torch.empty(2, dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        _native.expand_as(src.data_ptr(), src.shape[0], b.shape[0], b.shape[1], True, None, out.data_ptr(),
                          scratch.data_ptr(), torch.cuda.current_stream(a.device).cuda_stream)
    return (out >= 1).float()


def compute_iou(a: torch.Tensor, b: torch.Tensor) -> float:
    a = _match_size(a, b)
    intersection = (a * b).sum()
    union = a.sum() + b.sum() - intersection
    return (intersection / (union + 1e-8)).item()


def compute_ioa(a: torch.Tensor, b: torch.Tensor) -> float:
    a = _match_size(a, b)
    intersection = (a * b).sum()
    return (intersection / (a.sum() + 1e-8)).item()
