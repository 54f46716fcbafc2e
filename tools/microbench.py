#!/usr/bin/env python
"""Kernel-variant micro-benchmark (run on the GPU box): accumulate paths x RMW modes x launch granularity.

Prints one line per variant: ms per step, GB/s of algorithmic bytes, fraction of the measured HBM peak."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import algorithmic_bytes_per_step, build_sets, measured_peak, traced_layers  # noqa: E402
from daam_b200 import _native, ops  # noqa: E402


def time_variant(sets, flags, per_layer, steps=200, warmup=20):
    stream = torch.cuda.current_stream()
    n = len(sets)
    singles = [[ops.pack([s[0].array[j]]) for j in range(s[0].n)] for s in sets] if per_layer else None

    def step(i):
        if per_layer:
            for d in singles[i % n]:
                ops.accumulate(d, 'cuda', stream, flags)
        else:
            ops.accumulate(sets[i % n][0], 'cuda', stream, flags)

    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(int(8e6))       # ~4 ms gate: the launches below queue up behind it, so host pacing is not timed
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='sd21')
    ap.add_argument('--prompts', type=int, nargs='+', default=[1, 8])
    ap.add_argument('--dtypes', nargs='+', default=['bf16', 'fp32'])   # fp32: mma-* = split (2 x tf32) form
    ap.add_argument('--variants', nargs='+', default=None)
    ap.add_argument('--no-save', action='store_true', help='do not write gpurun_out/microbench_*.json (runs under ncu)')
    args = ap.parse_args()
    peak, _ = measured_peak()
    layers = traced_layers(args.workload)
    variants = {
        'simt-ldst': _native.ACC_FORCE_SIMT | _native.ACC_RMW_LDST,
        'simt-red': _native.ACC_FORCE_SIMT | _native.ACC_RMW_RED,
        'mma-red': _native.ACC_FORCE_MMA | _native.ACC_RMW_RED,
        'mma-ldst': _native.ACC_FORCE_MMA | _native.ACC_RMW_LDST,
        'mma-red-nopdl': _native.ACC_FORCE_MMA | _native.ACC_RMW_RED | _native.ACC_NO_PDL,
        'mma-red-early': _native.ACC_FORCE_MMA | _native.ACC_RMW_RED | _native.ACC_EARLY_LOADS,
        'mma-ldst-early': _native.ACC_FORCE_MMA | _native.ACC_RMW_LDST | _native.ACC_EARLY_LOADS,
    }
    if args.variants:
        variants = {k: v for k, v in variants.items() if k in args.variants}
    rows = []
    for dt in args.dtypes:
        dtype = {'bf16': torch.bfloat16, 'fp16': torch.float16, 'fp32': torch.float32}[dt]
        for p in args.prompts:
            nbytes = algorithmic_bytes_per_step(layers, p, 4 if dt == 'fp32' else 2)
            n_sets = max(2, -(-int(320e6) // (nbytes // 2)))
            sets = build_sets(layers, p, dtype, n_sets, 0)
            for name, flags in variants.items():
                for per_layer in (False, True):
                    try:
                        ms = time_variant(sets, flags, per_layer)
                    except _native.NativeError as e:
                        print(f'{args.workload} {dt} P={p} {name} per_layer={per_layer}: {e}', flush=True)
                        continue
                    gbs = nbytes / (ms * 1e-3) / 1e9
                    row = dict(workload=args.workload, dtype=dt, prompts=p, variant=name, per_layer=per_layer,
                               ms_per_step=round(ms, 5), gbs=round(gbs, 1), frac=round(gbs / peak, 4))
                    rows.append(row)
                    print(json.dumps(row), flush=True)
            del sets
            torch.cuda.empty_cache()
    if args.no_save:
        return
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', f'microbench_{args.workload}_{"-".join(args.dtypes)}.json'), 'w') as f:
        json.dump(rows, f, indent=1)


if __name__ == '__main__':
    main()
