"""world_size-2 gloo test of the only multi-process logic on the path: prompt sharding + the final gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from daam_b200.distributed import gather_heat_maps, pad_heat_map, shard_prompts


def test_shard_prompts_partition():
    prompts = [f'p{i}' for i in range(11)]
    parts = [shard_prompts(prompts, r, 4) for r in range(4)]
    assert sorted(i for p in parts for i in p) == list(range(11))
    assert max(map(len, parts)) - min(map(len, parts)) <= 1
    assert parts[1] == [1, 5, 9]
    with pytest.raises(ValueError):
        shard_prompts(prompts, 4, 4)


def _fake_map(i, x=8):
    n_rows = 3 + i % 4
    return torch.full((n_rows, x, x), float(i + 1)) + torch.arange(n_rows).view(-1, 1, 1)


def _worker(rank, world, port, n_total, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        mine = [_fake_map(i) for i in shard_prompts(list(range(n_total)), rank, world)]
        out = gather_heat_maps(mine, n_total, 8, tokens=10)
        torch.save(out, os.path.join(out_dir, f'r{rank}.pt'))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n_total', [4, 5, 1])     # 1 < world: rank 1 owns no prompt and still joins the collective
def test_gather_heat_maps_world2(tmp_path, n_total):
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, n_total, str(tmp_path)), nprocs=2, join=True)
    expect = torch.stack([pad_heat_map(_fake_map(i), 10) for i in range(n_total)])
    for r in range(2):
        got = torch.load(os.path.join(tmp_path, f'r{r}.pt'))
        assert got.shape == (n_total, 10, 8, 8)
        assert torch.equal(got, expect)


def test_gather_single_process():
    maps = [_fake_map(i) for i in range(3)]
    out = gather_heat_maps(maps, 3, 8, tokens=10)
    assert torch.equal(out, torch.stack([pad_heat_map(m, 10) for m in maps]))
