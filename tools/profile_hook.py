#!/usr/bin/env python
"""Where the eager hook overhead goes (run on the GPU box): wall time of un-hooked vs hooked forwards of the full-cost
SD-2.1 UNet, with the host time spent inside the tracer's _enqueue / flush, for the default step launch (on the forward's
stream), the side-stream variant (launch='overlap'), and with the whole flush stubbed out."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from daam_b200 import trace  # noqa: E402
from daam_b200.testing.synthetic import SD21_SPEC, make_pipeline  # noqa: E402


def main():
    dtype = torch.bfloat16
    pipe = make_pipeline(SD21_SPEC, body='full', dtype=dtype, device='cuda', seed=0, init_on_device=True)
    spec = pipe.unet.spec
    lat = torch.randn(2, spec.in_channels, spec.sample_size, spec.sample_size, device='cuda', dtype=dtype)
    emb = torch.randn(2, spec.tokens, spec.cross_attention_dim, device='cuda', dtype=dtype)
    t_dev = torch.full((1,), 500.0, device='cuda')
    n = 40

    def run(k):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(k):
            pipe.unet(lat, t_dev, emb)
        host = time.perf_counter() - t0          # host time to ISSUE k forwards
        torch.cuda.synchronize()
        return host / k * 1e3, (time.perf_counter() - t0) / k * 1e3

    res = {}
    with torch.no_grad():
        run(5)
        res['unhooked_issue_ms'], res['unhooked_wall_ms'] = run(n)
        for mode in ('step', 'overlap', 'no_flush'):
            with trace(pipe, launch='overlap' if mode == 'overlap' else 'step') as tc:
                acc = {'enqueue': 0.0, 'flush': 0.0, 'n_enq': 0, 'n_flush': 0}
                enq, fl = tc._enqueue, tc.flush

                def enqueue(*a, _enq=enq, **kw):
                    t = time.perf_counter()
                    r = _enq(*a, **kw)
                    acc['enqueue'] += time.perf_counter() - t
                    acc['n_enq'] += 1
                    return r

                def flush(_fl=fl):
                    t = time.perf_counter()
                    if mode == 'no_flush':
                        tc._refs = []
                        tc._n_pending = 0
                        tc._step_id += 1
                    else:
                        _fl()
                    acc['flush'] += time.perf_counter() - t
                    acc['n_flush'] += 1

                tc._enqueue, tc.flush = enqueue, flush
                run(5)
                acc.update(enqueue=0.0, flush=0.0, n_enq=0, n_flush=0)
                issue, wall = run(n)
                res[mode] = {'issue_ms': round(issue, 4), 'wall_ms': round(wall, 4),
                             'enqueue_us_per_step': round(acc['enqueue'] / n * 1e6, 1),
                             'flush_us_per_step': round(acc['flush'] / n * 1e6, 1),
                             'enqueue_calls_per_step': acc['n_enq'] / n, 'flush_calls_per_step': acc['n_flush'] / n}
                tc._n_pending = 0
                tc._refs = []
        res['unhooked2_issue_ms'], res['unhooked2_wall_ms'] = run(n)
    print(json.dumps(res, indent=1))
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'profile_hook.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
