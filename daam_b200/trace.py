"""The tracer: ``with trace(pipe) as tc: pipe(prompt); tc.compute_global_heat_map()`` on B200-native kernels.

Mirror of the reference's L1 (``/root/reference/daam/trace.py``): same classes, constructor arguments, methods and
exceptions; what differs is what runs underneath.

* The attention processor (:class:`UNetCrossAttentionHooker`, reference trace.py:189-315) owns the whole attn2 forward
  like the reference's, but never materialises the probabilities: the layer output comes from SDPA, and the
  heat-map side of the call -- ``get_attention_scores`` + ``_unravel_attn`` + the per-head ``update`` loop (trace.py:
  276, 219-244, 293-294) -- is one fused CUDA kernel (``daam_accumulate``) reading the Q/K projections in place.
* Kernel work is queued per denoising step and issued as ONE persistent launch covering every traced layer of the step
  at the end of the UNet forward, on the forward's own stream (``launch='step'``, default: one CUDA call per step, the
  kernel is 0.2 % of a step) or on a side stream so that it also overlaps the next step's first kernels
  (``launch='overlap'``: four CUDA calls per step); ``launch='layer'`` issues it immediately per layer.
* ``compute_global_heat_map`` (trace.py:83-132) keeps the Python-side key filter and error messages and runs the
  bicubic-upsample / clamp / mean / normalise reduction as one kernel (``daam_finalize``).

Accumulators are fp32 regardless of the pipeline dtype (the reference accumulates in the pipeline dtype, SURVEY.md
section 5); parity is stated against the fp32 oracle fed the same Q/K.
"""
from __future__ import annotations

import math
from pathlib import Path
from typing import Dict, List, Optional, Type, Union

import numpy as np
import torch
import torch.nn.functional as F

from . import _native, ops
from .heatmap import GlobalHeatMap, LayerSlab, RawHeatMapCollection
from .hook import AggregateHooker, ObjectHooker, UNetCrossAttentionLocator
from .utils import cache_dir

__all__ = ['trace', 'DiffusionHeatMapHooker', 'GlobalHeatMap', 'UNetCrossAttentionHooker', 'PipelineHooker',
           'ImageProcessorHooker']

class DiffusionHeatMapHooker(AggregateHooker):
    """Context manager that traces every located cross-attention layer of ``pipeline.unet`` (trace.py:22-59).

    Extra keyword-only options (not in the reference): ``launch`` ('step' | 'overlap' | 'layer', see module docstring),
    ``batch_prompts`` (accept several prompts per generation: N independent single-prompt traces sharing each launch;
    the reference rejects this, trace.py:172-173) and ``locate_middle_block`` (also locate the mid block without
    enabling save/load of heads -- BASELINE config 5 "all 16+70 layers").
    """

    def __init__(self, pipeline, low_memory: bool = False, load_heads: bool = False, save_heads: bool = False,
                 data_dir: str = None, *, launch: str = 'step', batch_prompts: bool = False,
                 locate_middle_block: bool = False, kernel_flags: int = _native.ACC_AUTO):
        if launch not in ('step', 'overlap', 'layer'):
            raise ValueError("launch must be 'step', 'overlap' or 'layer'")
        _native.load()   # fail here, loudly, if the CUDA library is missing
        self.all_heat_maps = RawHeatMapCollection()
        side = pipeline.unet.config.sample_size * pipeline.vae_scale_factor
        self.latent_hw = 4096 if side in (512, 1024) else 9216   # 64x64, or 96x96 for the 768-pixel models
        self.locator = UNetCrossAttentionLocator(restrict={0} if low_memory else None,
                                                 locate_middle_block=locate_middle_block or load_heads or save_heads)
        self.last_prompt: str = ''
        self.last_prompts: List[str] = []
        self.last_image = None
        self.time_idx = 0
        self._gen_idx = 0
        self.launch = launch
        self.batch_prompts = batch_prompts
        self.kernel_flags = kernel_flags
        # step queue: the layer calls of the running UNet forward, kept as one reusable host-side daam_layer[] whose slots
        # are rewritten in place (in the steady state a layer only stores two pointers into its slot)
        self._packed = _native.PackedLayers([_native.DaamLayer()] * 64)
        self._slots = [self._packed.array[i] for i in range(64)]      # ctypes proxies into the array, created once
        self._n_pending = 0
        self._refs: List[torch.Tensor] = []    # the queued projections, kept alive until their launch has run
        self._parked: List[list] = []          # projections of launches that may still be running
        self._layer_state: Dict[int, tuple] = {}   # layer -> (q shape, dtype, position, slot, slab, q_off, k_off, device, own)
        self._queued: Dict[int, int] = {}      # layer -> id of the step it was last queued in
        self._step_id = 1
        self._epoch_seen = -1                  # RawHeatMapCollection.epoch the cached layer states belong to
        self._device = None
        self._launcher: Optional[_native.SideLauncher] = None
        self._stream: Optional[torch.cuda.Stream] = None
        self._dirty = False                    # side-stream work not yet ordered before the current stream
        self.all_heat_maps.bind(self.synchronize, self._zero_slabs)

        modules = [
            UNetCrossAttentionHooker(m, self, layer_idx=idx, latent_hw=self.latent_hw, load_heads=load_heads,
                                     save_heads=save_heads, data_dir=data_dir)
            for idx, m in enumerate(self.locator.locate(pipeline.unet))
        ]
        modules.append(PipelineHooker(pipeline, self))
        if type(pipeline).__name__ == 'StableDiffusionXLPipeline' and getattr(pipeline, 'image_processor', None):
            modules.append(ImageProcessorHooker(pipeline.image_processor, self))
        super().__init__(modules)
        self.pipe = pipeline

    # -- small reference API ------------------------------------------------------------------------------------------
    def time_callback(self, *args, **kwargs):
        self.time_idx += 1

    @property
    def layer_names(self):
        return self.locator.layer_names

    def to_experiment(self, path, seed=None, id='.', subtype='.', **compute_kwargs):
        """Exports the last generation call to a serializable generation experiment (trace.py:68-81)."""
        from .experiment import GenerationExperiment
        return GenerationExperiment(
            self.last_image,
            self.compute_global_heat_map(**compute_kwargs).heat_maps,
            self.last_prompt,
            seed=seed, id=id, subtype=subtype, path=path, tokenizer=self.pipe.tokenizer,
        )

    def _hook_impl(self):
        super()._hook_impl()
        unet = self.pipe.unet
        self._forward_hook = None
        if self.launch != 'layer' and hasattr(unet, 'register_forward_hook'):
            # end of every UNet forward = end of the step's layer calls: issue the step launch right away
            self._forward_hook = unet.register_forward_hook(lambda *_: self.flush())

    def _unhook_impl(self):
        self.synchronize()      # issue what is queued and order it (and the parked projections) before the caller's stream
        if getattr(self, '_forward_hook', None) is not None:
            self._forward_hook.remove()
            self._forward_hook = None
        super()._unhook_impl()

    # -- kernel queue -------------------------------------------------------------------------------------------------
    def _side_stream(self, device) -> torch.cuda.Stream:
        if self._stream is None or self._stream.device != torch.device(device):
            self._stream = torch.cuda.Stream(device=device)
        return self._stream

    def _enqueue(self, layer_idx: int, factor: int, q: torch.Tensor, k: torch.Tensor, heads: int, scale: float):
        """Register one traced layer call: ``q [B, hw, C]``, ``k [B, 77, C]`` straight from ``to_q`` / ``to_k``.

        This runs once per layer per step on the host's critical path, so the steady state does as little as possible:
        when the projections are contiguous, shaped like the layer's previous call and the layer arrives at the same
        position of the step as last time, only the two data pointers of its slot in the step's ``daam_layer[]`` change."""
        if self._queued.get(layer_idx) == self._step_id:   # the layer comes round again: a new UNet forward has started
            self.flush()
        heat_maps = self.all_heat_maps
        if heat_maps.epoch != self._epoch_seen:            # a slab was (re)allocated: cached descriptors may be stale
            self._layer_state.clear()
            self._epoch_seen = heat_maps.epoch
        pos = self._n_pending if self.launch != 'layer' else 0
        st = self._layer_state.get(layer_idx)
        if st is not None and q.shape == st[0] and q.dtype is st[1] and st[2] == pos and q.is_contiguous() \
                and k.is_contiguous() and q.get_device() == st[7]:
            slot, slab = st[3], st[4]
            slot.q = q.data_ptr() + st[5]
            slot.k = k.data_ptr() + st[6]
            if not slab.touched:
                heat_maps.mark_live(slab)
        else:
            st, q, k = self._describe(layer_idx, factor, q, k, heads, scale, pos)    # (q, k: contiguous copies if it made any)
            slab = st[4]
        if self.launch == 'layer':
            if torch.cuda.is_current_stream_capturing():
                slab.captured = True
            self._launch_now(st[8], q.device)
            return
        self._refs.append(q)                               # kept alive until the step's launch has run
        self._refs.append(k)
        self._queued[layer_idx] = self._step_id
        self._n_pending = pos + 1

    def _describe(self, layer_idx: int, factor: int, q: torch.Tensor, k: torch.Tensor, heads: int, scale: float, pos: int):
        """Slow path of :meth:`_enqueue`: (re)build the layer's slab and descriptor and cache them. Returns the state and
        the tensors the descriptor points into (the caller keeps those alive, not the originals)."""
        if not q.is_cuda:
            raise RuntimeError('daam_b200 traces pipelines that live on a CUDA device only (there is no CPU '
                               'fallback)')
        if q.stride(-1) != 1 or k.stride(-1) != 1:
            q, k = q.contiguous(), k.contiguous()
        bsz, hw, _ = q.shape
        side = int(math.sqrt(hw))
        if side * side != hw:
            raise RuntimeError(f'layer {layer_idx}: {hw} query positions are not a square map')
        # "second half of the batch*heads axis" (trace.py:240): the conditional samples of a CFG batch
        _, n_samples, head0, n_heads = ops.cond_half(bsz, heads)
        # Samples vs prompts: diffusers repeats every prompt num_images_per_prompt times (prompt-major), and the
        # reference's keys then enumerate images x heads of its single prompt (the "head" index of a key runs over
        # the whole kept axis, trace.py:240, 293-294). The slab is therefore [prompts][images * heads]; the kernel
        # sees the same memory as [samples][heads].
        n_real = len(self.last_prompts) if self.last_prompts else (n_samples if self.batch_prompts else 1)
        if n_real > 1 and not self.batch_prompts:
            raise ValueError('Only single prompt generation is supported for heat map computation.')
        if n_samples % n_real != 0:
            raise RuntimeError(f'layer {layer_idx}: {n_samples} conditional samples for {n_real} prompts')
        images = n_samples // n_real
        slab = self.all_heat_maps.slab_for(layer_idx, factor, n_real, images * n_heads, side, side, q.device, head0)
        self._epoch_seen = self.all_heat_maps.epoch        # (this call may have bumped it; the other layers' slabs stand)
        desc = ops.make_layer_desc(q, k, slab.acc.view(n_samples, n_heads, slab.acc.shape[2], hw), heads, scale)
        if self.launch != 'layer':
            if pos >= len(self._slots):                    # grow the step array (SDXL: 70 layers)
                grown = _native.PackedLayers([_native.DaamLayer()] * (2 * len(self._slots)))
                for i in range(pos):
                    grown.array[i] = self._packed.array[i]
                self._packed = grown
                self._slots = [grown.array[i] for i in range(len(grown.array))]
                self._layer_state.clear()                  # cached slot proxies point into the old array
            self._packed.array[pos] = desc
            slot, own = self._slots[pos], None
        else:
            own = _native.PackedLayers([desc])             # 'layer' mode: every layer launches its own 1-element array
            slot = own.array[0]
        # the fast path is only valid for contiguous projections (strides implied by the shape)
        shape_key = q.shape if (q.is_contiguous() and k.is_contiguous()) else None
        state = (shape_key, q.dtype, pos, slot, slab, desc.q - q.data_ptr(), desc.k - k.data_ptr(), q.get_device(), own)
        self._layer_state[layer_idx] = state
        self._device = q.device
        return state, q, k

    def _launch_now(self, own, device):
        """``launch='layer'``: the layer's kernel right away on the current stream (the producer of Q/K may be the
        immediately preceding kernel there, so no EARLY_LOADS)."""
        index = device.index if device.index is not None else torch.cuda.current_device()
        stream = torch.cuda.current_stream(index).cuda_stream
        if index == torch.cuda.current_device():
            _native.accumulate(own, stream, self.kernel_flags)
        else:
            with torch.cuda.device(index):
                _native.accumulate(own, stream, self.kernel_flags)

    def _accumulate_probs(self, layer_idx: int, factor: int, probs: torch.Tensor, bsz: int, heads: int):
        """Heat maps from materialised probabilities (save_heads / load_heads compatibility path)."""
        hw = probs.shape[1]
        side = int(math.sqrt(hw))
        _, n_samples, head0, n_heads = ops.cond_half(bsz, heads)
        n_real = len(self.last_prompts) if self.last_prompts else (n_samples if self.batch_prompts else 1)
        if n_real > 1 and not self.batch_prompts:
            raise ValueError('Only single prompt generation is supported for heat map computation.')
        if n_samples % n_real != 0:
            raise RuntimeError(f'layer {layer_idx}: {n_samples} conditional samples for {n_real} prompts')
        slab = self.all_heat_maps.slab_for(layer_idx, factor, n_real, (n_samples // n_real) * n_heads, side, side,
                                           probs.device, head0)
        self.synchronize()
        ops.accumulate_probs(probs, slab.acc)

    def flush(self):
        """Issue the queued layer calls as one persistent launch (per pack of 32 layers)."""
        n = self._n_pending
        if n == 0:
            return
        device = self._device
        index = device.index if device.index is not None else torch.cuda.current_device()
        packed = self._packed
        packed.n = n
        flags = self.kernel_flags | _native.ACC_EARLY_LOADS
        current = torch.cuda.current_stream(index).cuda_stream
        previous = torch.cuda.current_device()
        switch = index != previous
        if switch:
            torch.cuda.set_device(index)
        try:
            capturing = torch.cuda.is_current_stream_capturing()
            if capturing:
                # CUDA-graph capture of the UNet step: the launch becomes a node of the captured stream; replays bypass the
                # Python hook, so the layers of this launch stay live across per-generation resets
                step = self._step_id
                for layer_idx, st in self._layer_state.items():
                    if self._queued.get(layer_idx) == step:
                        st[4].captured = True
            if capturing or self.launch == 'step':
                # On the forward's own stream: the predecessor there is the tail of the UNet forward, never a producer of
                # the queued Q/K, so only the accumulator updates have to wait for it (EARLY_LOADS). Stream order also
                # makes it safe to drop the projections right after the launch.
                _native.accumulate(packed, current, flags)
            else:
                side = self._side_stream(device)
                if self._launcher is None:
                    self._launcher = _native.SideLauncher()
                if self._parked and self._launcher.idle():     # the previous launches have run: drop their projections
                    self._parked = []
                elif len(self._parked) >= 8:                   # the host runs many steps ahead of the device: do not let
                    side.synchronize()                         # parked projections pile up (they pin allocator blocks)
                    self._parked = []
                # One foreign call: event on the current stream (Q/K were produced there) -> the side stream waits ->
                # launch -> `done` event. The side stream carries nothing but these launches and the projections are
                # complete before the previous one could have started (EARLY_LOADS holds here too).
                self._launcher.launch(packed, flags, current, side.cuda_stream)
                self._parked.append(self._refs)                # alive until a later idle() / join says the kernel has run
                self._dirty = True
        finally:
            if switch:
                torch.cuda.set_device(previous)
        self._refs = []
        self._n_pending = 0
        self._step_id += 1

    def synchronize(self):
        """Make every accumulate issued so far visible to work enqueued on the current stream afterwards."""
        self.flush()
        if self._dirty and self._stream is not None:
            self._launcher.join(torch.cuda.current_stream(self._stream.device).cuda_stream)
            self._dirty = False
            # parked projections: their memory may only be reused by the current stream after the side stream is done
            # with them, which the wait above now guarantees for everything issued so far
            self._parked = []

    def _zero_slabs(self, slabs: List[LayerSlab]):
        if not slabs:
            return
        self.synchronize()
        for slab in slabs:
            slab.acc.zero_()
        if self._stream is not None:   # later side-stream launches must see the zeroed slabs
            self._stream.wait_stream(torch.cuda.current_stream(slabs[0].acc.device))

    # -- finalize -------------------------------------------------------------------------------------------------------
    def compute_global_heat_map(self, prompt=None, factors=None, head_idx=None, layer_idx=None, normalize=False,
                                prompt_idx: int = 0) -> GlobalHeatMap:
        """Aggregate across time (already summed in the slabs) and across layers/heads (trace.py:83-132).

        Args mirror the reference: ``factors`` restricts the spatial factors, ``head_idx`` / ``layer_idx`` restrict to one
        head / layer, ``normalize`` divides by the per-pixel sum over the real tokens. ``prompt_idx`` selects the prompt in
        ``batch_prompts`` mode.
        """
        if prompt is None:
            prompt = self.last_prompts[prompt_idx] if self.last_prompts else self.last_prompt
        factors = {0, 1, 2, 4, 8, 16, 32, 64} if factors is None else set(factors)
        x = int(np.sqrt(self.latent_hw))
        self.synchronize()
        groups, keep = [], []
        for slab in self.all_heat_maps.live_slabs():
            if slab.factor not in factors or (layer_idx is not None and layer_idx != slab.layer_idx):
                continue
            if head_idx is not None and not 0 <= head_idx < slab.heads:
                continue
            acc = slab.acc[prompt_idx]
            groups.append(_native.DaamKeyGroup(acc=acc.data_ptr(), heads=slab.heads, h=slab.h, w=slab.w,
                                               tokens=acc.shape[1], head_sel=-1 if head_idx is None else head_idx,
                                               reserved=0))
            keep.append(acc)
        if not groups:
            if head_idx is not None or layer_idx is not None:
                raise RuntimeError('No heat maps found for the given parameters.')
            raise RuntimeError('No heat maps found. Did you forget to call `with trace(...)` during generation?')
        n_rows = min(len(self.pipe.tokenizer.tokenize(prompt)) + 2, _native.TOKENS)   # 1 for SOS and 1 for padding
        device = keep[0].device
        maps = torch.empty((n_rows, x, x), dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            _native.finalize(groups, x, n_rows, normalize, maps.data_ptr(),
                             torch.cuda.current_stream(device).cuda_stream)
        return GlobalHeatMap(self.pipe.tokenizer, prompt, maps)


    def compute_per_head_heat_maps(self, prompt=None, factors=None, normalize=False, prompt_idx: int = 0):
        """Every ``compute_global_heat_map(layer_idx=l, head_idx=h)`` of the reference's ``--all-heads`` sweep
        (daam/run/generate.py:239-255) in one launch. Returns ``(keys, maps)``: ``keys[i] = (factor, layer, head)`` and
        ``maps[i]`` the ``[n_tokens + 2, x, x]`` heat map the reference computes for that single key."""
        if prompt is None:
            prompt = self.last_prompts[prompt_idx] if self.last_prompts else self.last_prompt
        factors = {0, 1, 2, 4, 8, 16, 32, 64} if factors is None else set(factors)
        x = int(np.sqrt(self.latent_hw))
        self.synchronize()
        groups, keep, keys = [], [], []
        for slab in self.all_heat_maps.live_slabs():
            if slab.factor not in factors:
                continue
            acc = slab.acc[prompt_idx]
            groups.append(_native.DaamKeyGroup(acc=acc.data_ptr(), heads=slab.heads, h=slab.h, w=slab.w,
                                               tokens=acc.shape[1], head_sel=-1, reserved=0))
            keep.append(acc)
            keys += [(slab.factor, slab.layer_idx, head) for head in range(slab.heads)]
        if not groups:
            raise RuntimeError('No heat maps found. Did you forget to call `with trace(...)` during generation?')
        n_rows = min(len(self.pipe.tokenizer.tokenize(prompt)) + 2, _native.TOKENS)
        device = keep[0].device
        maps = torch.empty((len(keys), n_rows, x, x), dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            _native.finalize_per_key(groups, x, n_rows, normalize, maps.data_ptr(),
                                     torch.cuda.current_stream(device).cuda_stream)
        return keys, maps


class ImageProcessorHooker(ObjectHooker):
    """Remembers the first post-processed image of an SDXL pipeline (trace.py:135-147)."""

    def __init__(self, processor, parent_trace: 'trace'):
        super().__init__(processor)
        self.parent_trace = parent_trace

    def _hooked_postprocess(hk_self, _, *args, **kwargs):
        images = hk_self.monkey_super('postprocess', *args, **kwargs)
        hk_self.parent_trace.last_image = images[0]
        return images

    def _hook_impl(self):
        self.monkey_patch('postprocess', self._hooked_postprocess)


class PipelineHooker(ObjectHooker):
    """Per-generation reset + prompt capture at ``check_inputs``; image capture at the safety checker (trace.py:150-186)."""

    def __init__(self, pipeline, parent_trace: 'trace'):
        super().__init__(pipeline)
        self.heat_maps = parent_trace.all_heat_maps
        self.parent_trace = parent_trace

    def _hooked_run_safety_checker(hk_self, self, image, *args, **kwargs):
        image, has_nsfw = hk_self.monkey_super('run_safety_checker', image, *args, **kwargs)
        processor = getattr(self, 'image_processor', None)
        if processor:
            images = processor.postprocess(image, output_type='pil') if torch.is_tensor(image) \
                else processor.numpy_to_pil(image)
        else:
            images = self.numpy_to_pil(image)
        hk_self.parent_trace.last_image = images[len(images) - 1]
        return image, has_nsfw

    def _hooked_check_inputs(hk_self, _, prompt: Union[str, List[str]], *args, **kwargs):
        tr = hk_self.parent_trace
        if isinstance(prompt, str):
            prompts = [prompt]
        else:
            prompts = list(prompt)
            if len(prompts) > 1 and not tr.batch_prompts:
                raise ValueError('Only single prompt generation is supported for heat map computation.')
        hk_self.heat_maps.clear()
        if len(prompts) != len(tr.last_prompts):    # slabs are laid out [prompts][images * heads]: re-derive them
            tr._layer_state.clear()
        tr.last_prompt = prompts[0]
        tr.last_prompts = prompts
        return hk_self.monkey_super('check_inputs', prompt, *args, **kwargs)

    def _hook_impl(self):
        self.monkey_patch('run_safety_checker', self._hooked_run_safety_checker, strict=False)  # absent in SDXL
        self.monkey_patch('check_inputs', self._hooked_check_inputs)


class UNetCrossAttentionHooker(ObjectHooker):
    """The attention processor installed on one ``attn2`` module (trace.py:189-315)."""

    def __init__(self, module, parent_trace: 'trace', context_size: int = 77, layer_idx: int = 0,
                 latent_hw: int = 9216, load_heads: bool = False, save_heads: bool = False,
                 data_dir: Union[str, Path] = None):
        super().__init__(module)
        self.heat_maps = parent_trace.all_heat_maps
        self.context_size = context_size
        self.layer_idx = layer_idx
        self.latent_hw = latent_hw
        self.load_heads = load_heads
        self.save_heads = save_heads
        self.trace = parent_trace
        self._geom = None
        self.data_dir = Path(data_dir) if data_dir is not None else cache_dir() / 'heads'
        if load_heads or save_heads:
            self.data_dir.mkdir(parents=True, exist_ok=True)

    def _save_attn(self, attn_slice: torch.Tensor):
        torch.save(attn_slice, self.data_dir / f'{self.trace._gen_idx}.pt')

    def _load_attn(self) -> torch.Tensor:
        return torch.load(self.data_dir / f'{self.trace._gen_idx}.pt')

    def _materialised_call(self, attn, query, key, value):
        """save_heads / load_heads (trace.py:276-302): the probabilities exist as a tensor -- written to / replaced
        from ``data_dir/{gen_idx}.pt`` -- heat maps and the layer output are both computed from that tensor."""
        bsz, n, _ = query.shape
        heads, tokens = attn.heads, key.shape[1]
        if self.save_heads:
            probs = ops.attention_probs(query, key, heads, attn.scale)     # [B*H, hw, 77], dtype of the pipeline
            self._save_attn(probs)
        else:
            probs = self._load_attn().to(query.device)
        factor = int(math.sqrt(self.latent_hw // probs.shape[1]))
        self.trace._gen_idx += 1
        if probs.shape[-1] == self.context_size and factor != 8:
            self.trace._accumulate_probs(self.layer_idx, factor, probs, bsz, heads)
        d = value.shape[-1] // heads
        v = value.view(bsz, tokens, heads, d).permute(0, 2, 1, 3).reshape(bsz * heads, tokens, d)
        out = torch.bmm(probs.to(v.dtype), v)
        out = out.view(bsz, heads, n, d).permute(0, 2, 1, 3).reshape(bsz, n, heads * d)
        return attn.to_out[1](attn.to_out[0](out))

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None):
        """attn2 forward: projections -> (heat-map kernel on Q/K) -> SDPA -> output projection."""
        if attention_mask is not None:
            raise RuntimeError('the heat-map kernel does not take an attention mask (SD cross-attention passes none)')
        bsz, n, _ = hidden_states.shape
        query = attn.to_q(hidden_states)
        if encoder_hidden_states is None:
            encoder_hidden_states = hidden_states
        elif attn.norm_cross is not None:
            encoder_hidden_states = attn.norm_cross(encoder_hidden_states)
        key = attn.to_k(encoder_hidden_states)
        value = attn.to_v(encoder_hidden_states)

        if self.save_heads or self.load_heads:
            return self._materialised_call(attn, query, key, value)

        heads = attn.heads
        tokens = key.shape[1]
        geom = self._geom                                    # (n, tokens) -> factor and the trace / skip decision
        if geom is None or geom[0] != n or geom[1] != tokens:
            factor = int(math.sqrt(self.latent_hw // n))
            geom = self._geom = (n, tokens, factor, tokens == self.context_size and factor != 8)   # trace.py:285-289
        tr = self.trace
        tr._gen_idx += 1
        if geom[3]:                                          # skip if too large (trace.py:289)
            tr._enqueue(self.layer_idx, geom[2], query, key, heads, attn.scale)

        d = query.shape[-1] // heads
        q4 = query.view(bsz, n, heads, d).transpose(1, 2)
        k4 = key.view(bsz, tokens, heads, d).transpose(1, 2)
        v4 = value.view(bsz, tokens, heads, d).transpose(1, 2)
        out = F.scaled_dot_product_attention(q4, k4, v4, scale=attn.scale)
        out = out.transpose(1, 2).reshape(bsz, n, heads * d)
        out = attn.to_out[0](out)    # linear proj
        return attn.to_out[1](out)   # dropout

    def _hook_impl(self):
        self.original_processor = self.module.processor
        self.module.set_processor(self)

    def _unhook_impl(self):
        self.module.set_processor(self.original_processor)

    @property
    def num_heat_maps(self):
        return len(self.heat_maps)


trace: Type[DiffusionHeatMapHooker] = DiffusionHeatMapHooker
