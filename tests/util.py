"""Shared helpers of the test-suite (oracle access, fixture loading, CPU-side recomputation of GPU inputs)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
LAYER_FIXTURES = ['layer_hw256_h2_d64', 'layer_hw1024_h1_d64_peaky', 'layer_hw64_h2_d40', 'layer_hw576_h1_d64']


def golden(name):
    return np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)


def oracle_layer_maps(q, k, heads, scale, steps=1):
    """Oracle rows a3+a4 (+a6 summed over `steps` identical calls) for q [B, hw, C], k [B, 77, C]: [N*H, 77, hw] fp32.

    The inputs are moved to CPU fp32 first: products of fp16/bf16 values are exact in fp32, so the oracle sees exactly
    the values the kernel reads (parity is against the fp32 oracle fed identical Q/K)."""
    from oracle import daam_oracle as O
    maps = O.port_layer_step(q.detach().float().cpu(), k.detach().float().cpu(), heads, scale)
    maps = maps.reshape(maps.shape[0], maps.shape[1], -1)
    return maps * steps if steps != 1 else maps


def rel_err(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()
