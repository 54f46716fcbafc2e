#!/usr/bin/env python
"""Finalize-family micro-benchmark (run on the GPU box): daam_finalize over the traced keys of a workload with a cold L2
(rotating slab sets), the per-key sweep, and the fused word-list expansion vs the per-word loop."""
import argparse
import json
import os
import sys
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import traced_layers  # noqa: E402
from daam_b200 import _native  # noqa: E402
from daam_b200.heatmap import GlobalHeatMap  # noqa: E402
from daam_b200.testing.synthetic import WhitespaceTokenizer  # noqa: E402


def timed(fn, n, gate=True):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if gate:
        torch.cuda._sleep(int(4e6))
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3     # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='sd21')
    ap.add_argument('--rows', type=int, default=16)
    ap.add_argument('--no-save', action='store_true', help='do not write gpurun_out/microbench_finalize_*.json (runs under ncu)')
    args = ap.parse_args()
    layers = traced_layers(args.workload)
    x = int(max(hw for hw, _, _ in layers) ** 0.5)        # 64, or 96 for the 768-pixel geometry
    slab_bytes = sum(h * 77 * hw * 4 for hw, h, _ in layers)
    n_sets = max(2, -(-int(400e6) // slab_bytes))
    sets = []
    for s in range(n_sets):
        slabs = [torch.rand(h, 77, hw, device='cuda') for hw, h, _ in layers]
        groups = [_native.DaamKeyGroup(acc=t.data_ptr(), heads=t.shape[0], h=int(t.shape[2] ** 0.5), w=int(t.shape[2] ** 0.5),
                                       tokens=77, head_sel=-1, reserved=0) for t in slabs]
        sets.append((slabs, groups))
    out = torch.empty(args.rows, x, x, device='cuda')
    stream = torch.cuda.current_stream().cuda_stream
    n_keys = sum(h for _, h, _ in layers)
    res = {'workload': args.workload, 'keys': n_keys, 'rows': args.rows, 'slab_mb': slab_bytes / 1e6,
           'read_mb': slab_bytes * args.rows / 77 / 1e6}
    for normalize in (False, True):
        f = lambda i: _native.finalize(sets[i % n_sets][1], x, args.rows, normalize, out.data_ptr(), stream)
        for i in range(3):
            f(i)
        res[f'finalize_us_norm{int(normalize)}'] = round(timed(f, 20), 2)
    os.environ['DAAM_FINALIZE_GENERIC'] = '1'
    f = lambda i: _native.finalize(sets[i % n_sets][1], x, args.rows, False, out.data_ptr(), stream)
    f(0)
    res['finalize_generic_us'] = round(timed(f, 10), 2)
    os.environ['DAAM_FINALIZE_GENERIC'] = '0'
    res['finalize_gbs'] = round(res['read_mb'] / res['finalize_us_norm0'] * 1e3, 1)
    # fused word list vs the per-word loop (compute_word_heat_map + expand_as, 4 launches + a blocking D2H per word)
    prompt = 'a photo of a dog chasing a red ball on the beach'
    g = torch.rand(len(prompt.split()) + 2, x, x, device='cuda')
    ghm = GlobalHeatMap(WhitespaceTokenizer(), prompt, g)
    img = SimpleNamespace(size=(512, 512))
    words = prompt.split()
    for _ in range(2):
        ghm.expand_words(words, img)
        [ghm.compute_word_heat_map(w).expand_as(img) for w in words]
    import time
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        ghm.expand_words(words, img)
    torch.cuda.synchronize()
    res['expand_words_fused_us_per_call_wall'] = round((time.perf_counter() - t0) / 20 * 1e6, 1)
    t0 = time.perf_counter()
    for _ in range(20):
        [ghm.compute_word_heat_map(w).expand_as(img) for w in words]
    torch.cuda.synchronize()
    res['expand_per_word_loop_us_per_call_wall'] = round((time.perf_counter() - t0) / 20 * 1e6, 1)
    res['expand_words_device_us'] = round(timed(lambda i: ghm.expand_words(words, img, to_cpu=False), 20, gate=False), 2)
    res['n_words'] = len(words)
    print(json.dumps(res))
    if args.no_save:
        return
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', f'microbench_finalize_{args.workload}.json'), 'w') as f:
        json.dump(res, f, indent=1)


if __name__ == '__main__':
    main()
