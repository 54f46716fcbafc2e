"""Heat-map store and algebra: the accumulator slabs the CUDA kernel sums into, and the global / word heat maps.

Mirror of the reference's L2 (``/root/reference/daam/heatmap.py``) for the hot-path rows of SURVEY.md section 8a:

* :class:`RawHeatMapCollection` (heatmap.py:148-172) -- same interface (``update``, ``factors``, ``layers``, ``heads``,
  iteration over ``((factor, layer, head), tensor[77, h, w])``, ``clear``), but backed by one fp32 device slab per traced
  layer, laid out ``[prompts][heads][77][h*w]``: the slab *is* the reference's per-key tensors (each key a contiguous
  view), and it is what ``daam_accumulate`` adds into, so nothing is copied or re-laid-out between kernel and API.
* :class:`GlobalHeatMap` (heatmap.py:114-142) / :class:`WordHeatMap` (heatmap.py:56-96) -- ``compute_word_heat_map``
  and ``expand_as`` run the native kernels (``daam_word_heat_map``, ``daam_expand_as``).

The spaCy-parsed iterators and matplotlib overlays of the reference are out of scope;
``plot_overlay`` is kept as a thin optional-matplotlib helper because ``expand_as(plot=True)`` calls it.
"""
from __future__ import annotations

from dataclasses import dataclass
from functools import lru_cache
from typing import Dict, Iterator, List, Optional, Set, Tuple

import torch

from . import _native
from .utils import compute_token_merge_indices

__all__ = ['GlobalHeatMap', 'RawHeatMapCollection', 'WordHeatMap', 'LayerSlab']

RawHeatMapKey = Tuple[int, int, int]  # factor, layer, head


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f'{what}: daam_b200 computes on CUDA tensors only (there is no CPU fallback)')


@dataclass
class LayerSlab:
    """Accumulators of one traced layer: ``acc[prompt][head]`` is the reference's ``[77, h, w]`` map of key
    ``(factor, layer_idx, head)``."""
    layer_idx: int
    factor: int
    heads: int
    h: int
    w: int
    acc: torch.Tensor            # fp32 [n_prompts, heads, 77, h*w]
    touched: bool = False        # a key exists only once it has been updated (defaultdict semantics)
    head_offset: int = 0         # first real head behind key head 0 (non-zero only for the un-guided B=1 quirk)
    captured: bool = False       # the layer's kernel launch is part of a CUDA graph: replays update it without the hook

    @property
    def n_prompts(self) -> int:
        return self.acc.shape[0]

    def key_view(self, head: int, prompt: int = 0) -> torch.Tensor:
        return self.acc[prompt, head].view(self.acc.shape[2], self.h, self.w)


class RawHeatMapCollection:
    """Per-(factor, layer, head) time-sums of attention maps, resident in HBM as fp32 slabs."""

    def __init__(self):
        self.slabs: Dict[int, LayerSlab] = {}
        self._order: List[int] = []          # layer indices in first-update order (the reference's dict order)
        self.epoch = 0                        # bumped whenever a slab object is (re)allocated (descriptor caches key on it)
        self._sync = None                     # callable making pending kernel work visible to the current stream
        self._zero = None                     # callable(slabs) zeroing slabs in accumulate-stream order

    # -- wiring from the tracer -------------------------------------------------------------------------------------
    def bind(self, sync, zero):
        self._sync, self._zero = sync, zero

    def _synchronize(self):
        if self._sync is not None:
            self._sync()

    def slab_for(self, layer_idx: int, factor: int, n_prompts: int, heads: int, h: int, w: int, device,
                 head_offset: int = 0) -> LayerSlab:
        """Returns (allocating or re-shaping on demand) the zero-initialised slab of a layer and marks it live."""
        slab = self.slabs.get(layer_idx)
        shape = (n_prompts, heads, _native.TOKENS, h * w)
        if slab is None or tuple(slab.acc.shape) != shape or slab.acc.device != torch.device(device) \
                or slab.factor != factor:
            if torch.cuda.is_available() and torch.cuda.is_current_stream_capturing():
                raise RuntimeError('accumulator slabs cannot be created inside a CUDA-graph capture: run one eager '
                                   'UNet step under trace() before capturing')
            acc = torch.zeros(shape, dtype=torch.float32, device=device)
            slab = LayerSlab(layer_idx, factor, heads, h, w, acc, head_offset=head_offset)
            self.slabs[layer_idx] = slab
            self.epoch += 1
        if not slab.touched:
            slab.touched = True
            self._order.append(layer_idx)
        return slab

    def mark_live(self, slab: LayerSlab):
        if not slab.touched:
            slab.touched = True
            self._order.append(slab.layer_idx)

    # -- reference interface ------------------------------------------------------------------------------------------
    def update(self, factor: int, layer_idx: int, head_idx: int, heatmap: torch.Tensor):
        """``acc[key] += heatmap`` for an externally produced ``[77, h, w]`` map (heatmap.py:153-156). The traced path
        never calls this -- the kernel accumulates in place -- it exists for API compatibility (e.g. merging maps)."""
        _require_cuda(heatmap, 'RawHeatMapCollection.update')
        self._synchronize()
        t, h, w = heatmap.shape
        slab = self.slabs.get(layer_idx)
        heads = max(head_idx + 1, slab.heads if slab is not None and slab.touched else 0)
        if slab is None or not slab.touched or slab.heads < heads or (slab.h, slab.w) != (h, w):
            old = slab if slab is not None and slab.touched and (slab.h, slab.w) == (h, w) else None
            acc = torch.zeros((1, heads, t, h * w), dtype=torch.float32, device=heatmap.device)
            if old is not None:
                acc[:, :old.heads] = old.acc[:1]
            new = LayerSlab(layer_idx, factor, heads, h, w, acc, touched=True)
            self.slabs[layer_idx] = new
            self.epoch += 1
            if layer_idx not in self._order:
                self._order.append(layer_idx)
            slab = new
        slab.acc[0, head_idx] += heatmap.reshape(t, h * w).float()

    def live_slabs(self) -> List[LayerSlab]:
        return [self.slabs[i] for i in self._order]

    def factors(self) -> Set[int]:
        return {s.factor for s in self.live_slabs()}

    def layers(self) -> Set[int]:
        return {s.layer_idx for s in self.live_slabs()}

    def heads(self) -> Set[int]:
        return {h for s in self.live_slabs() for h in range(s.heads)}

    def items(self, prompt: int = 0) -> Iterator[Tuple[RawHeatMapKey, torch.Tensor]]:
        self._synchronize()
        for slab in self.live_slabs():
            for head in range(slab.heads):
                yield (slab.factor, slab.layer_idx, head), slab.key_view(head, prompt)

    def __iter__(self):
        return self.items(0)

    def __len__(self):
        return sum(s.heads for s in self.live_slabs())

    def clear(self):
        """Forget every key (heatmap.py:170-172). Slabs stay allocated and are zeroed in stream order for re-use."""
        live = self.live_slabs()
        if self._zero is not None:
            self._zero(live)
        else:
            for slab in live:
                slab.acc.zero_()
        for slab in live:                     # graph-captured layers stay live: replays bypass the Python hook
            slab.touched = slab.captured
        self._order = [i for i in self._order if self.slabs[i].captured]


class WordHeatMap:
    def __init__(self, heatmap: torch.Tensor, word: str = None, word_idx: int = None):
        self.word = word
        self.word_idx = word_idx
        self.heatmap = heatmap

    @property
    def value(self):
        return self.heatmap

    def expand_as(self, image, absolute: bool = False, threshold: Optional[float] = None, plot: bool = False,
                  **plot_kwargs) -> torch.Tensor:
        """Bicubic-upsample to the image size, min-max normalise unless ``absolute``, optionally binarise; returns a
        CPU tensor like heatmap.py:77-93 (including its ``size=(image.size[0], image.size[1])`` axis order)."""
        _require_cuda(self.heatmap, 'WordHeatMap.expand_as')
        src = self.heatmap.detach().float().contiguous()
        out_h, out_w = int(image.size[0]), int(image.size[1])
        out = torch.empty((out_h, out_w), dtype=torch.float32, device=src.device)
        scratch = torch.empty(_native.EXPAND_SCRATCH_FLOATS, dtype=torch.float32, device=src.device)
        with torch.cuda.device(src.device):
            _native.expand_as(src.data_ptr(), src.shape[-1], out_h, out_w, absolute, threshold, out.data_ptr(),
                              scratch.data_ptr(), _stream_ptr(src.device))
        im = out.cpu()          # the reference returns a CPU tensor (heatmap.py:93); GlobalHeatMap.expand_words defers this
        if plot:
            self.plot_overlay(image, **plot_kwargs)
        return im

    def plot_overlay(self, image, out_file=None, color_normalize=True, ax=None, **expand_kwargs):
        """Optional visual helper (heatmap.py:20-53, 66-75); needs matplotlib, which the hot path does not."""
        try:
            from matplotlib import pyplot as plt
        except ImportError as e:  # pragma: no cover
            raise RuntimeError('plot_overlay needs matplotlib, which is not part of the heat-map hot path') from e
        import numpy as np
        heat = self.expand_as(image, **expand_kwargs)
        target = plt if ax is None else ax
        if color_normalize:
            target.imshow(heat.numpy(), cmap='jet')
        else:
            heat = heat.clamp(0, 1)
            target.imshow(heat.numpy(), cmap='jet', vmin=0.0, vmax=1.0)
        im = torch.from_numpy(np.array(image)).float() / 255
        target.imshow(torch.cat((im, 1 - heat.unsqueeze(-1)), dim=-1))
        if self.word is not None:
            (plt.title if ax is None else ax.set_title)(self.word)
        if out_file is not None:
            plt.savefig(out_file)

    def compute_ioa(self, other: 'WordHeatMap') -> float:
        """Intersection over own area -- daam/evaluate.py:26-35 (heatmap.py:95-96)."""
        from .evaluate import compute_ioa
        return compute_ioa(self.heatmap, other.heatmap)


class GlobalHeatMap:
    """``[n_prompt_tokens + 2, x, x]`` per-token maps plus the word lookup (heatmap.py:114-123)."""

    def __init__(self, tokenizer, prompt: str, heat_maps: torch.Tensor):
        self.tokenizer = tokenizer
        self.heat_maps = heat_maps
        self.prompt = prompt
        self.compute_word_heat_map = lru_cache(maxsize=50)(self.compute_word_heat_map)

    def compute_word_heat_map(self, word: str, word_idx: int = None, offset_idx: int = 0) -> WordHeatMap:
        rows, word_idx = compute_token_merge_indices(self.tokenizer, self.prompt, word, word_idx, offset_idx)
        maps = self.heat_maps
        _require_cuda(maps, 'GlobalHeatMap.compute_word_heat_map')
        n_rows, x = maps.shape[0], maps.shape[-1]
        for r in rows:  # torch's advanced indexing raises IndexError on out-of-range rows
            if not -n_rows <= r < n_rows:
                raise IndexError(f'index {r} is out of bounds for dimension 0 with size {n_rows}')
        maps = maps.detach().float().contiguous()
        out = torch.empty((x, x), dtype=torch.float32, device=maps.device)
        with torch.cuda.device(maps.device):
            _native.word_heat_map(maps.data_ptr(), n_rows, x, rows, out.data_ptr(), _stream_ptr(maps.device))
        return WordHeatMap(out, word, word_idx)

    def expand_words(self, words, image, absolute: bool = False, threshold: Optional[float] = None,
                     word_idx=None, offset_idx: int = 0, to_cpu: bool = True):
        """``[self.compute_word_heat_map(w).expand_as(image, absolute, threshold) for w in words]`` -- the loop every user
        of the reference writes (heatmap.py:121-123 then 77-93) -- as ONE fused launch (gather-mean of the word's rows ->
        bicubic to the image size -> min/max -> normalise / threshold) and one device-to-host copy instead of four
        launches and a blocking copy per word.

        Returns ``(word_heat_maps, expanded)``: a list of :class:`WordHeatMap` (device ``[x, x]`` views, same values as
        ``compute_word_heat_map``) and ``expanded`` ``[len(words), image.size[0], image.size[1]]`` (CPU by default like
        ``expand_as``; ``to_cpu=False`` keeps it on the device until the caller needs it). ``word_idx`` may be a list
        parallel to ``words``. Raises the reference's ``ValueError`` for a word that is not in the prompt."""
        words = list(words)
        idxs = list(word_idx) if isinstance(word_idx, (list, tuple)) else [word_idx] * len(words)
        merged = [compute_token_merge_indices(self.tokenizer, self.prompt, w, i, offset_idx) for w, i in zip(words, idxs)]
        maps = self.heat_maps
        _require_cuda(maps, 'GlobalHeatMap.expand_words')
        n_rows, x = maps.shape[0], maps.shape[-1]
        for rows, _ in merged:
            for r in rows:
                if not -n_rows <= r < n_rows:
                    raise IndexError(f'index {r} is out of bounds for dimension 0 with size {n_rows}')
        if not words:
            return [], torch.empty((0, int(image.size[0]), int(image.size[1])))
        maps = maps.detach().float().contiguous()
        out_h, out_w = int(image.size[0]), int(image.size[1])
        dev = maps.device
        word_maps = torch.empty((len(words), x, x), dtype=torch.float32, device=dev)
        out = torch.empty((len(words), out_h, out_w), dtype=torch.float32, device=dev)
        scratch = torch.empty(_native.EXPAND_SCRATCH_FLOATS * len(words), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _native.expand_words(maps.data_ptr(), n_rows, x, [rows for rows, _ in merged], out_h, out_w, absolute,
                                 threshold, word_maps.data_ptr(), out.data_ptr(), scratch.data_ptr(), _stream_ptr(dev))
        whms = [WordHeatMap(word_maps[i], w, idx) for i, (w, (_, idx)) in enumerate(zip(words, merged))]
        return whms, (out.cpu() if to_cpu else out)
