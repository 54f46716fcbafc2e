"""Synthetic, random-init stand-ins for the diffusers objects the hot path plugs into.

``diffusers`` is not installed on either box and there is no network, so the benchmark and the tests run
on plain-torch modules that expose exactly the surface the reference touches
(``/root/reference/daam/trace.py:252-311``, ``/root/reference/daam/hook.py:95-127``):

* :class:`SyntheticAttention` -- the ``diffusers==0.21.2`` ``Attention`` module surface: ``to_q/to_k/to_v/to_out``,
  ``heads``, ``scale``, ``norm_cross``, ``upcast_attention``, ``upcast_softmax``, ``processor``/``set_processor`` and the
  helper methods the reference's hook calls (``prepare_attention_mask``, ``head_to_batch_dim``, ``batch_to_head_dim``,
  ``get_attention_scores``); SURVEY.md section 8c spells out the 0.21.2 semantics restated here.
* :class:`SyntheticUNet` -- a UNet2DConditionModel-shaped module tree (``down_blocks``/``mid_block``/``up_blocks`` whose
  class names contain ``CrossAttn``, ``.attentions[*].transformer_blocks[*].attn2``, ``config.sample_size``) in the
  SD-2.1-base and SDXL shapes. ``body='skeleton'`` keeps only the cross-attention layers (cheap; the operator-boundary
  benchmark and the CPU oracle use it), ``body='full'`` adds the resnets / self-attention / feed-forward so that
  "hooked vs un-hooked forward" is a meaningful overhead measurement.
* :class:`SyntheticPipeline` -- ``unet``, ``vae_scale_factor``, ``tokenizer.tokenize``, ``check_inputs``,
  ``image_processor.postprocess`` and a classifier-free-guidance loop that feeds the UNet ``[uncond, cond]`` batches.

These are fixtures: they contain no DAAM logic. The weights are random (default torch init), the data synthetic.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from types import SimpleNamespace
from typing import List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

__all__ = [
    'SyntheticAttention', 'SDPAProcessor', 'SyntheticUNet', 'SyntheticPipeline', 'WhitespaceTokenizer',
    'UNetSpec', 'SD21_SPEC', 'SD21_768_SPEC', 'SDXL_SPEC', 'SD15_SPEC', 'TINY_SPEC', 'TINY15_SPEC', 'TINY96_SPEC', 'make_pipeline',
]


# ---------------------------------------------------------------------------------------------------------------
# Attention module with the diffusers 0.21.2 surface
# ---------------------------------------------------------------------------------------------------------------
class SDPAProcessor:
    """The un-hooked baseline: what diffusers' ``AttnProcessor2_0`` does (projections -> SDPA -> out projection)."""

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None):
        b, n, _ = hidden_states.shape
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        if encoder_hidden_states is not None and attn.norm_cross is not None:
            ctx = attn.norm_cross(ctx)
        h = attn.heads
        q = attn.to_q(hidden_states).view(b, n, h, -1).transpose(1, 2)
        k = attn.to_k(ctx).view(b, ctx.shape[1], h, -1).transpose(1, 2)
        v = attn.to_v(ctx).view(b, ctx.shape[1], h, -1).transpose(1, 2)
        out = F.scaled_dot_product_attention(q, k, v, attn_mask=attention_mask)
        out = out.transpose(1, 2).reshape(b, n, -1)
        return attn.to_out[1](attn.to_out[0](out))


class SyntheticAttention(nn.Module):
    """``diffusers.models.attention_processor.Attention`` (0.21.2) restated: same attributes, same helper semantics."""

    def __init__(self, query_dim: int, cross_attention_dim: Optional[int] = None, heads: int = 8, dim_head: int = 64,
                 upcast_attention: bool = False, upcast_softmax: bool = False):
        super().__init__()
        inner = heads * dim_head
        ctx_dim = query_dim if cross_attention_dim is None else cross_attention_dim
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.upcast_attention = upcast_attention
        self.upcast_softmax = upcast_softmax
        self.norm_cross = None
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(ctx_dim, inner, bias=False)
        self.to_v = nn.Linear(ctx_dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])
        self.processor = SDPAProcessor()

    def set_processor(self, processor):
        self.processor = processor

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask)

    # -- helpers the reference hook calls (trace.py:261, 272-276, 297) --------------------------------------
    def prepare_attention_mask(self, attention_mask, target_length, batch_size):
        if attention_mask is None:
            return None
        if attention_mask.shape[-1] != target_length:
            attention_mask = F.pad(attention_mask, (0, target_length), value=0.0)
        if attention_mask.shape[0] < batch_size * self.heads:
            attention_mask = attention_mask.repeat_interleave(self.heads, dim=0)
        return attention_mask

    def head_to_batch_dim(self, tensor):
        b, n, c = tensor.shape
        h = self.heads
        return tensor.reshape(b, n, h, c // h).permute(0, 2, 1, 3).reshape(b * h, n, c // h)

    def batch_to_head_dim(self, tensor):
        bh, n, d = tensor.shape
        h = self.heads
        return tensor.reshape(bh // h, h, n, d).permute(0, 2, 1, 3).reshape(bh // h, n, d * h)

    def get_attention_scores(self, query, key, attention_mask=None):
        dtype = query.dtype
        if self.upcast_attention:
            query, key = query.float(), key.float()
        if attention_mask is None:
            base = torch.empty(query.shape[0], query.shape[1], key.shape[1], dtype=query.dtype, device=query.device)
            beta = 0
        else:
            base, beta = attention_mask, 1
        scores = torch.baddbmm(base, query, key.transpose(-1, -2), beta=beta, alpha=self.scale)
        del base
        if self.upcast_softmax:
            scores = scores.float()
        probs = scores.softmax(dim=-1)
        del scores
        return probs.to(dtype)


# ---------------------------------------------------------------------------------------------------------------
# UNet pieces
# ---------------------------------------------------------------------------------------------------------------
class _GEGLUFeedForward(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.proj = nn.Linear(dim, dim * 8)
        self.out = nn.Linear(dim * 4, dim)

    def forward(self, x):
        a, g = self.proj(x).chunk(2, dim=-1)
        return self.out(a * F.gelu(g))


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, dim_head, ctx_dim, full: bool, upcast_attention: bool = False):
        super().__init__()
        self.full = full
        if full:
            self.norm1 = nn.LayerNorm(dim)
            self.attn1 = SyntheticAttention(dim, None, heads, dim_head, upcast_attention=upcast_attention)
            self.norm3 = nn.LayerNorm(dim)
            self.ff = _GEGLUFeedForward(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = SyntheticAttention(dim, ctx_dim, heads, dim_head, upcast_attention=upcast_attention)

    def forward(self, x, ctx):
        if self.full:
            x = x + self.attn1(self.norm1(x))
        x = x + self.attn2(self.norm2(x), encoder_hidden_states=ctx)
        if self.full:
            x = x + self.ff(self.norm3(x))
        return x


class Transformer2DModel(nn.Module):
    def __init__(self, channels, heads, dim_head, ctx_dim, depth, full, upcast_attention=False):
        super().__init__()
        self.norm = nn.GroupNorm(32 if channels % 32 == 0 else 1, channels, eps=1e-6)
        self.proj_in = nn.Linear(channels, channels)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(channels, heads, dim_head, ctx_dim, full, upcast_attention) for _ in range(depth)])
        self.proj_out = nn.Linear(channels, channels)

    def forward(self, x, ctx):
        b, c, h, w = x.shape
        res = x
        y = self.norm(x).permute(0, 2, 3, 1).reshape(b, h * w, c)
        y = self.proj_in(y)
        for blk in self.transformer_blocks:
            y = blk(y, ctx)
        y = self.proj_out(y).reshape(b, h, w, c).permute(0, 3, 1, 2)
        return y + res


class ResnetBlock2D(nn.Module):
    def __init__(self, cin, cout, temb_dim, full):
        super().__init__()
        self.full = full
        if full:
            self.norm1 = nn.GroupNorm(32 if cin % 32 == 0 else 1, cin, eps=1e-5)
            self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
            self.time_emb_proj = nn.Linear(temb_dim, cout)
            self.norm2 = nn.GroupNorm(32 if cout % 32 == 0 else 1, cout, eps=1e-5)
            self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.shortcut = nn.Conv2d(cin, cout, 1) if (cin != cout or not full) else None

    def forward(self, x, temb):
        if not self.full:
            return self.shortcut(x)
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        return h + (x if self.shortcut is None else self.shortcut(x))


class _Down(nn.Module):
    def __init__(self, c, full):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=1) if full else None

    def forward(self, x):
        return self.conv(x) if self.conv is not None else F.avg_pool2d(x, 2)


class _Up(nn.Module):
    def __init__(self, c, full):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1) if full else None

    def forward(self, x):
        x = F.interpolate(x, scale_factor=2.0, mode='nearest')
        return self.conv(x) if self.conv is not None else x


class _DownBlockBase(nn.Module):
    def __init__(self, cin, cout, temb_dim, n_layers, add_down, full, attn=None):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, temb_dim, full)
                                      for i in range(n_layers)])
        if attn is not None:
            self.attentions = nn.ModuleList([Transformer2DModel(cout, full=full, **attn) for _ in range(n_layers)])
        self.downsamplers = nn.ModuleList([_Down(cout, full)]) if add_down else None

    def forward(self, x, temb, ctx):
        skips = []
        for i, res in enumerate(self.resnets):
            x = res(x, temb)
            if hasattr(self, 'attentions'):
                x = self.attentions[i](x, ctx)
            skips.append(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            skips.append(x)
        return x, skips


class DownBlock2D(_DownBlockBase):
    pass


class CrossAttnDownBlock2D(_DownBlockBase):
    pass


class _UpBlockBase(nn.Module):
    def __init__(self, cin_prev, cout, skip_channels: Sequence[int], temb_dim, add_up, full, attn=None):
        super().__init__()
        res = []
        for i, sc in enumerate(skip_channels):
            res.append(ResnetBlock2D((cin_prev if i == 0 else cout) + sc, cout, temb_dim, full))
        self.resnets = nn.ModuleList(res)
        if attn is not None:
            self.attentions = nn.ModuleList(
                [Transformer2DModel(cout, full=full, **attn) for _ in range(len(skip_channels))])
        self.upsamplers = nn.ModuleList([_Up(cout, full)]) if add_up else None

    def forward(self, x, skips, temb, ctx):
        for i, res in enumerate(self.resnets):
            x = res(torch.cat([x, skips.pop()], dim=1), temb)
            if hasattr(self, 'attentions'):
                x = self.attentions[i](x, ctx)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class UpBlock2D(_UpBlockBase):
    pass


class CrossAttnUpBlock2D(_UpBlockBase):
    pass


class UNetMidBlock2DCrossAttn(nn.Module):
    def __init__(self, c, temb_dim, full, attn):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, temb_dim, full), ResnetBlock2D(c, c, temb_dim, full)])
        self.attentions = nn.ModuleList([Transformer2DModel(c, full=full, **attn)])

    def forward(self, x, temb, ctx):
        x = self.resnets[0](x, temb)
        x = self.attentions[0](x, ctx)
        return self.resnets[1](x, temb)


@dataclass
class UNetSpec:
    """Shape of a UNet2DConditionModel. ``heads[i]``/``depth[i]`` belong to ``block_out_channels[i]``; ``depth`` 0
    means a block without cross-attention (``DownBlock2D``/``UpBlock2D``)."""
    name: str
    sample_size: int
    block_out_channels: Sequence[int]
    heads: Sequence[int]
    depth: Sequence[int]
    cross_attention_dim: int
    dim_head: Optional[int] = 64              # None: SD-1.x style, head_dim = channels // heads
    layers_per_block: int = 2
    mid_depth: Optional[int] = None          # transformer depth of the mid block (None: same as last block, min 1)
    in_channels: int = 4
    upcast_attention: bool = False
    tokens: int = 77


# public unet/config.json values of stabilityai/stable-diffusion-2-1-base and stabilityai/stable-diffusion-xl-base-1.0
SD21_SPEC = UNetSpec('sd21-base', 64, (320, 640, 1280, 1280), (5, 10, 20, 20), (1, 1, 1, 0), 1024)
# stabilityai/stable-diffusion-2-1 (the 768-pixel v-prediction model): same UNet, 96x96 latent -> latent_hw 9216
SD21_768_SPEC = UNetSpec('sd21-768', 96, (320, 640, 1280, 1280), (5, 10, 20, 20), (1, 1, 1, 0), 1024)
SDXL_SPEC = UNetSpec('sdxl-base', 128, (320, 640, 1280), (5, 10, 20), (0, 2, 10), 2048, mid_depth=10)
# runwayml/stable-diffusion-v1-5: 8 heads at every level, head dims 40 / 80 / 160
SD15_SPEC = UNetSpec('sd15', 64, (320, 640, 1280, 1280), (8, 8, 8, 8), (1, 1, 1, 0), 768, dim_head=None)
# a small tree with the SD-2.1 topology (15 located layers, factors 1/2/4, mid layer at factor 8) for CPU tests
TINY_SPEC = UNetSpec('tiny', 64, (64, 128, 128, 128), (1, 2, 2, 2), (1, 1, 1, 0), 96)
# the same topology with SD-1.x style head dims (40 / 80 / 80)
TINY15_SPEC = UNetSpec('tiny15', 64, (80, 160, 160, 160), (2, 2, 2, 2), (1, 1, 1, 0), 96, dim_head=None)
# the 768-pixel models' geometry (96x96 latent -> latent_hw 9216, daam/trace.py:32-33): layers at 96^2 / 48^2 / 24^2
# (9216 / 2304 / 576 query positions: partial 128-pixel tiles), mid layer at 12^2 (factor 8, skipped)
TINY96_SPEC = UNetSpec('tiny96', 96, (64, 128, 128, 128), (1, 2, 2, 2), (1, 1, 1, 0), 96)


class SyntheticUNet(nn.Module):
    """UNet2DConditionModel-shaped random-init network (see module docstring). ``forward(sample, t, ctx)``."""

    def __init__(self, spec: UNetSpec, body: str = 'skeleton'):
        super().__init__()
        assert body in ('skeleton', 'full')
        full = body == 'full'
        self.spec = spec
        self.config = SimpleNamespace(sample_size=spec.sample_size, in_channels=spec.in_channels,
                                      cross_attention_dim=spec.cross_attention_dim)
        ch = list(spec.block_out_channels)
        temb_dim = ch[0] * 4
        self.temb_dim = temb_dim
        self.time_embedding = nn.Sequential(nn.Linear(ch[0], temb_dim), nn.SiLU(), nn.Linear(temb_dim, temb_dim))
        self.conv_in = nn.Conv2d(spec.in_channels, ch[0], 3, padding=1)

        def attn_cfg(i):
            if spec.depth[i] == 0:
                return None
            return dict(heads=spec.heads[i], dim_head=spec.dim_head or ch[i] // spec.heads[i],
                        ctx_dim=spec.cross_attention_dim, depth=spec.depth[i], upcast_attention=spec.upcast_attention)

        downs, skip_ch = [], [ch[0]]
        cin = ch[0]
        for i, cout in enumerate(ch):
            last = i == len(ch) - 1
            cls = CrossAttnDownBlock2D if spec.depth[i] else DownBlock2D
            downs.append(cls(cin, cout, temb_dim, spec.layers_per_block, not last, full, attn_cfg(i)))
            skip_ch += [cout] * spec.layers_per_block + ([] if last else [cout])
            cin = cout
        self.down_blocks = nn.ModuleList(downs)

        mid_depth = spec.mid_depth if spec.mid_depth is not None else max(1, spec.depth[-1])
        self.mid_block = UNetMidBlock2DCrossAttn(ch[-1], temb_dim, full, dict(
            heads=spec.heads[-1], dim_head=spec.dim_head or ch[-1] // spec.heads[-1],
            ctx_dim=spec.cross_attention_dim, depth=mid_depth,
            upcast_attention=spec.upcast_attention))

        ups = []
        cin = ch[-1]
        for j, i in enumerate(reversed(range(len(ch)))):
            cout = ch[i]
            last = j == len(ch) - 1
            sk = [skip_ch.pop() for _ in range(spec.layers_per_block + 1)]
            cls = CrossAttnUpBlock2D if spec.depth[i] else UpBlock2D
            ups.append(cls(cin, cout, sk, temb_dim, not last, full, attn_cfg(i)))
            cin = cout
        self.up_blocks = nn.ModuleList(ups)
        self.conv_norm_out = nn.GroupNorm(32 if ch[0] % 32 == 0 else 1, ch[0])
        self.conv_out = nn.Conv2d(ch[0], spec.in_channels, 3, padding=1)

    def _time_embed(self, t, batch, dtype, device):
        half = self.spec.block_out_channels[0] // 2
        freqs = torch.exp(-math.log(10000.0) * torch.arange(half, device=device, dtype=torch.float32) / half)
        arg = torch.as_tensor(t, device=device, dtype=torch.float32).reshape(-1, 1).expand(batch, 1) * freqs[None]
        emb = torch.cat([arg.sin(), arg.cos()], dim=-1).to(dtype)
        return self.time_embedding(emb)

    def forward(self, sample, timestep, encoder_hidden_states):
        temb = self._time_embed(timestep, sample.shape[0], sample.dtype, sample.device)
        x = self.conv_in(sample)
        skips = [x]
        for blk in self.down_blocks:
            x, s = blk(x, temb, encoder_hidden_states)
            skips += s
        x = self.mid_block(x, temb, encoder_hidden_states)
        for blk in self.up_blocks:
            x = blk(x, skips, temb, encoder_hidden_states)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


# ---------------------------------------------------------------------------------------------------------------
# Pipeline
# ---------------------------------------------------------------------------------------------------------------
class WhitespaceTokenizer:
    """CLIP-tokenizer stand-in: lower-cased whitespace pieces carrying the ``</w>`` end-of-word marker."""

    def tokenize(self, text: str) -> List[str]:
        return [w + '</w>' for w in text.lower().split()]


class _ImageProcessor:
    def postprocess(self, image, output_type='pil'):
        return [image[i] for i in range(image.shape[0])]


class SyntheticPipeline:
    """A StableDiffusionPipeline-shaped driver around :class:`SyntheticUNet`.

    Per denoising step it copies the step's inputs (latents, text embeddings, timestep; kept in pinned host memory when
    the UNet lives on a GPU) to the device, runs the UNet on the CFG batch ``[uncond x N, cond x N]`` and reads the
    guided noise estimate's mean back to the host -- the host<->device traffic ``bench.py`` counts for its ``e2e`` figure.

    ``cuda_graph=True`` replays the step's device work (UNet forward + guidance update) from a CUDA graph: the first step
    of a configuration runs eagerly, the second is captured, later steps and later calls replay it. Attention processors
    installed at capture time (e.g. a tracer's) are part of the graph; the graph is re-captured when they change.
    """

    def __init__(self, unet: SyntheticUNet, dtype=torch.float32, device='cpu', seed: int = 0,
                 cuda_graph: bool = False):
        self.unet = unet.to(device=device, dtype=dtype).eval()
        self.dtype, self.device = dtype, torch.device(device)
        self.vae_scale_factor = 8
        self.tokenizer = WhitespaceTokenizer()
        self.image_processor = _ImageProcessor()
        self.seed = seed
        self.cuda_graph = cuda_graph and self.device.type == 'cuda'
        self.h2d_bytes_per_step = 0
        self.d2h_bytes_per_step = 0
        self._graphs = {}
        self._attn = [m for m in self.unet.modules() if isinstance(m, SyntheticAttention)]

    def check_inputs(self, prompt, *args, **kwargs):
        if not isinstance(prompt, (str, list)):
            raise ValueError('`prompt` has to be of type `str` or `list`')

    def encode(self, prompts: List[str], generator: torch.Generator):
        """Synthetic text encoder: seeded gaussian embeddings, [uncond x N, cond x N] like diffusers' CFG concat."""
        n, spec = len(prompts), self.unet.spec
        emb = torch.randn(2 * n, spec.tokens, spec.cross_attention_dim, generator=generator, dtype=torch.float32)
        return emb

    def _step(self, st, guidance_scale):
        """Device work of one denoising step on the static buffers ``st``."""
        lat, n = st['latents'], st['latents'].shape[0]
        eps = self.unet(torch.cat([lat, lat], dim=0), st['t'], st['emb'])
        eps_u, eps_c = eps[:n], eps[n:]
        eps = eps_u + guidance_scale * (eps_c - eps_u)
        lat.copy_((lat - 0.02 * eps).clamp_(-4, 4))
        st['stat'].copy_(eps.float().mean(dim=(1, 2, 3)))

    def _state(self, n):
        spec, dev = self.unet.spec, self.device
        key = (n, tuple(id(m.processor) for m in self._attn))
        st = self._graphs.get(key)
        if st is None:
            if len(self._graphs) > 4:
                self._graphs.clear()
            st = {
                'emb': torch.empty(2 * n, spec.tokens, spec.cross_attention_dim, dtype=self.dtype, device=dev),
                'lat0': torch.empty(n, spec.in_channels, spec.sample_size, spec.sample_size, dtype=self.dtype, device=dev),
                'latents': torch.empty(n, spec.in_channels, spec.sample_size, spec.sample_size, dtype=self.dtype,
                                       device=dev),
                't': torch.zeros(1, dtype=torch.float32, device=dev),
                'stat': torch.zeros(n, dtype=torch.float32, device=dev),
                'graph': None, 'eager_steps': 0,
            }
            self._graphs[key] = st
        return st

    @torch.no_grad()
    def __call__(self, prompt, num_inference_steps: int = 50, generator: Optional[torch.Generator] = None,
                 callback=None, guidance_scale: float = 7.5):
        self.check_inputs(prompt)
        prompts = [prompt] if isinstance(prompt, str) else list(prompt)
        n, spec = len(prompts), self.unet.spec
        if generator is None:
            generator = torch.Generator().manual_seed(self.seed)
        cuda = self.device.type == 'cuda'
        emb_h = self.encode(prompts, generator).to(self.dtype)
        lat_h = torch.randn(n, spec.in_channels, spec.sample_size, spec.sample_size, generator=generator,
                            dtype=torch.float32).to(self.dtype)
        t_h = torch.tensor([[1000.0 * (1.0 - i / max(1, num_inference_steps))] for i in range(num_inference_steps)])
        out_h = torch.empty(n, dtype=torch.float32)
        if cuda:
            emb_h, lat_h, t_h, out_h = emb_h.pin_memory(), lat_h.pin_memory(), t_h.pin_memory(), out_h.pin_memory()
        self.h2d_bytes_per_step = emb_h.numel() * emb_h.element_size() + lat_h.numel() * lat_h.element_size() + 4
        self.d2h_bytes_per_step = n * 4
        st = self._state(n)
        for i in range(num_inference_steps):
            st['emb'].copy_(emb_h, non_blocking=True)            # H2D: the step's inputs
            st['lat0'].copy_(lat_h, non_blocking=True)
            st['t'].copy_(t_h[i], non_blocking=True)
            if i == 0:
                st['latents'].copy_(st['lat0'])
            if self.cuda_graph and st['graph'] is not None:
                st['graph'].replay()
            elif self.cuda_graph and st['eager_steps'] >= 1:
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    self._step(st, guidance_scale)
                st['graph'] = graph
                graph.replay()                                     # capture does not execute: run the step now
            else:
                self._step(st, guidance_scale)
                st['eager_steps'] += 1
            out_h.copy_(st['stat'], non_blocking=True)           # D2H: the step's result
            if callback is not None:
                callback(i, float(t_h[i]), st['latents'])
        latents = st['latents'].clone()
        image = latents[:, :3].float()
        images = self.image_processor.postprocess(image, output_type='pil')
        return SimpleNamespace(images=images, latents=latents)


def make_pipeline(spec: UNetSpec = SD21_SPEC, body: str = 'skeleton', dtype=torch.float32, device='cpu',
                  seed: int = 0, init_on_device: bool = False, cuda_graph: bool = False) -> SyntheticPipeline:
    """Random-init pipeline. Weights are drawn on the CPU from ``seed`` (identical on every box) unless
    ``init_on_device`` (fast for the full-size bodies; values then depend on the device RNG)."""
    gen_state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    if init_on_device and torch.device(device).type == 'cuda':
        with torch.device(device):
            unet = SyntheticUNet(spec, body=body)
    else:
        unet = SyntheticUNet(spec, body=body)
    torch.random.set_rng_state(gen_state)
    return SyntheticPipeline(unet, dtype=dtype, device=device, seed=seed, cuda_graph=cuda_graph)
