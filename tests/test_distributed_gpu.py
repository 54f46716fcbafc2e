"""Two ranks, two GPUs, NCCL: each rank traces its own prompts, one all_gather returns every finished map."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
PROMPTS = ['a red ball', 'two dogs on the beach', 'a cat', 'a dog chasing a red ball on the beach', 'a green tree']


def _trace_prompt(pipe, prompt, seed):
    from daam_b200 import trace
    with trace(pipe) as tc:
        pipe(prompt, num_inference_steps=2, generator=torch.Generator().manual_seed(seed))
        return tc.compute_global_heat_map().heat_maps.clone()


def _worker(rank, world, port, out_dir):
    from daam_b200.distributed import gather_heat_maps, shard_prompts
    from daam_b200.testing.synthetic import TINY_SPEC, make_pipeline
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        pipe = make_pipeline(TINY_SPEC, dtype=torch.float16, device=f'cuda:{rank}', seed=0)
        mine = [_trace_prompt(pipe, PROMPTS[i], 100 + i) for i in shard_prompts(PROMPTS, rank, world)]
        out = gather_heat_maps(mine, len(PROMPTS), 64)
        torch.save(out.cpu(), os.path.join(out_dir, f'r{rank}.pt'))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_sharded_prompts_and_nccl_gather(tmp_path):
    from daam_b200.distributed import pad_heat_map
    from daam_b200.testing.synthetic import TINY_SPEC, make_pipeline
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    pipe = make_pipeline(TINY_SPEC, dtype=torch.float16, device='cuda:0', seed=0)
    expect = torch.stack([pad_heat_map(_trace_prompt(pipe, p, 100 + i)) for i, p in enumerate(PROMPTS)]).cpu()
    for r in range(2):
        got = torch.load(os.path.join(tmp_path, f'r{r}.pt'))
        assert got.shape == (len(PROMPTS), 77, 64, 64)
        torch.testing.assert_close(got, expect, rtol=1e-4, atol=1e-5)   # same kernels, other GPU
