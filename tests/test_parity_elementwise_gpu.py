"""Element-wise parity (SURVEY.md section 8c's stated form) and the geometries / sweeps round 1 left to global bounds.

Tolerances, written here as the contract asks:

* per-key accumulators, fp32 inputs:  |got - ref| <= 1e-6 * steps + 1e-5 * |ref|   for every element;
* per-key accumulators, fp16 / bf16:  the oracle is fed the identical half-rounded Q/K in fp32 (products exact, only the
  summation order and ex2.approx differ): the same bound x 10;
* global heat maps:                   |got - ref| <= 1e-5 * steps + 1e-4 * |ref|   (x 10 for half inputs).

`rel_err` (max|err| / max|ref|) in the older tests bounds the large elements only; the small probabilities a DAAM user
looks at after normalisation are bounded here.
"""
import numpy as np
import pytest
import torch

from daam_b200 import _native, ops, trace
from daam_b200.testing.synthetic import TINY96_SPEC, TINY_SPEC, make_pipeline
from oracle import daam_oracle as O
from tests.test_finalize_gpu import fixture_groups, run_finalize
from tests.util import LAYER_FIXTURES, HookRecorder, assert_elementwise, golden, oracle_layer_maps

pytestmark = pytest.mark.gpu
DEV = 'cuda'
RTOL = {torch.float32: 1e-5, torch.float16: 1e-4, torch.bfloat16: 1e-4}
ATOL = {torch.float32: 1e-6, torch.float16: 1e-5, torch.bfloat16: 1e-5}       # x steps

PATHS = [('auto', _native.ACC_AUTO), ('simt', _native.ACC_FORCE_SIMT), ('mma', _native.ACC_FORCE_MMA),
         ('mma-early', _native.ACC_FORCE_MMA | _native.ACC_EARLY_LOADS)]


@pytest.fixture(autouse=True)
def _exact_fp32():
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize('path,flags', PATHS)
@pytest.mark.parametrize('name', LAYER_FIXTURES)
@pytest.mark.parametrize('dtype', [torch.float32, torch.float16])
def test_golden_layers_elementwise(name, dtype, path, flags):
    """The verbatim reference's maps (fixtures; inputs fp16-representable, so both dtypes read the reference's values)."""
    fx = golden(name)
    q = torch.from_numpy(fx['q']).to(DEV, dtype)
    k = torch.from_numpy(fx['k']).to(DEV, dtype)
    heads = int(fx['heads'])
    acc = ops.accumulate_layer(q, k, heads, float(fx['scale']), flags=flags)
    torch.cuda.synchronize()
    ref = torch.from_numpy(fx['maps']).reshape(1, heads, 77, -1)
    assert_elementwise(acc, ref, RTOL[dtype], ATOL[dtype], f'{name}/{path}/{dtype}')


SHAPES = [  # hw, heads, head_dim: SD-2.1 / SDXL layers, SD-1.x head dims, the 96-latent sizes incl. hw = 9216
    (4096, 5, 64), (1024, 10, 64), (256, 20, 64), (1024, 8, 80), (256, 8, 160), (4096, 8, 40),
    (9216, 5, 64), (2304, 10, 64), (576, 20, 64), (144, 20, 64), (9216, 8, 40), (64, 2, 192), (64, 2, 256),
]


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize('hw,heads,d', SHAPES)
def test_seeded_shapes_elementwise(hw, heads, d, dtype):
    g = torch.Generator().manual_seed(hw * 131 + heads * 7 + d)
    q = (torch.randn(2, hw, heads * d, generator=g) * 1.5).to(dtype).to(DEV)
    k = torch.randn(2, 77, heads * d, generator=g).to(dtype).to(DEV)
    ref = oracle_layer_maps(q, k, heads, d ** -0.5).unsqueeze(0)
    steps = 3
    for flags in (_native.ACC_AUTO, _native.ACC_FORCE_SIMT):
        acc = ops.new_accumulator(1, heads, hw, DEV)
        for _ in range(steps):
            ops.accumulate_layer(q, k, heads, acc=acc, flags=flags)
        torch.cuda.synchronize()
        assert_elementwise(acc, steps * ref, RTOL[dtype], ATOL[dtype] * steps, f'hw{hw} H{heads} d{d} {dtype} flags {flags}')


def test_tf32_split_error_budget():
    """fp32 projections on tensor cores (two tf32 terms per value, three products): how far from the fp32 oracle, and from
    a float64 statement of the same softmax. The budget is the fp32 contract (rtol 1e-5); the float64 column shows that
    the oracle's own fp32 rounding is of the same size as the kernel's."""
    g = torch.Generator().manual_seed(31)
    hw, heads, d = 1024, 4, 64
    for gain in (1.0, 3.0):
        q = (torch.randn(2, hw, heads * d, generator=g) * gain).to(DEV)
        k = torch.randn(2, 77, heads * d, generator=g).to(DEV)
        mma = ops.accumulate_layer(q, k, heads, flags=_native.ACC_FORCE_MMA)
        simt = ops.accumulate_layer(q, k, heads, flags=_native.ACC_FORCE_SIMT)
        torch.cuda.synchronize()
        q1 = q[1].cpu().double().reshape(hw, heads, d).permute(1, 0, 2).numpy()
        k1 = k[1].cpu().double().reshape(77, heads, d).permute(1, 0, 2).numpy()
        exact = torch.from_numpy(O.math_layer_maps(q1, k1, d ** -0.5)).reshape(1, heads, 77, hw)
        ref = oracle_layer_maps(q, k, heads, d ** -0.5).unsqueeze(0)
        for name, got in (('tf32-split', mma), ('simt', simt)):
            assert_elementwise(got, ref, 1e-5, 1e-6, f'{name} vs oracle, gain {gain}')
            assert_elementwise(got, exact, 1e-5, 1e-6, f'{name} vs float64, gain {gain}')


@pytest.mark.parametrize('case,select,head_sel,normalize', [
    ('global', lambda f, l, h: True, None, False),
    ('global_norm', lambda f, l, h: True, None, True),
    ('factors_2_4', lambda f, l, h: f in (2, 4), None, False),
    ('layer_1', lambda f, l, h: l == 1, None, False),
    ('head_1', lambda f, l, h: True, 1, False),
    ('layer_2_head_0', lambda f, l, h: l == 2, 0, False),
])
def test_finalize_golden_elementwise(case, select, head_sel, normalize):
    fx = golden('finalize')
    groups, keep = fixture_groups(fx, select, head_sel)
    ref = fx[case]
    out = run_finalize(groups, ref.shape[0], normalize)
    assert_elementwise(out, ref, 1e-4, 1e-5 * 2, case)          # the fixture's keys are sums over 2 steps


@pytest.mark.parametrize('normalize', [False, True])
def test_per_key_finalize_vs_reference_sweep(normalize):
    """daam_finalize_per_key against the reference's own --all-heads sweep (daam/run/generate.py:239-255): every
    compute_global_heat_map(layer_idx=l, head_idx=h) the verbatim reference produced (tests/golden/perkey.npz), and
    against the oracle port key by key."""
    fx, pk = golden('finalize'), golden('perkey')
    groups, keep = fixture_groups(fx)
    n_rows = pk['plain'].shape[1]
    want = pk['norm' if normalize else 'plain']
    out = torch.empty((want.shape[0], n_rows, 64, 64), device=DEV)
    _native.finalize_per_key(groups, 64, n_rows, normalize, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    # enumeration order of the kernel: group by group (layer order), head by head == the fixture's key order
    order = [(int(fx_key.split('_')[2]), int(fx_key.split('_')[3])) for fx_key in sorted(
        (n for n in fx.files if n.startswith('key_')), key=lambda n: (int(n.split('_')[2]), int(n.split('_')[3])))]
    assert [[l, h] for l, h in order] == [k[1:] for k in pk['keys'].tolist()]
    assert_elementwise(out, want, 1e-4, 2e-5, f'per-key sweep normalize={normalize}')
    keys = [((int(n.split('_')[1]), int(n.split('_')[2]), int(n.split('_')[3])), torch.from_numpy(fx[n]))
            for n in fx.files if n.startswith('key_')]
    for i, (l, h) in enumerate(order):
        ref = O.port_global_heat_map(keys, 4096, n_rows - 2, layer_idx=l, head_idx=h, normalize=normalize)
        assert_elementwise(out[i], ref, 1e-4, 2e-5, f'key layer {l} head {h}')


def test_per_head_maps_of_a_trace_vs_oracle():
    """compute_per_head_heat_maps on a traced generation: every (layer, head) map against the oracle's
    port_global_heat_map(layer_idx, head_idx) computed from the identical Q/K the hooks saw."""
    pipe = make_pipeline(TINY_SPEC, dtype=torch.float16, device=DEV, seed=8)
    prompt = 'a dog chasing a red ball on the beach'
    with trace(pipe) as tc:
        rec = HookRecorder(tc)
        pipe(prompt, num_inference_steps=2, generator=torch.Generator().manual_seed(1))
        store = rec.oracle_store()
        n_tok = len(pipe.tokenizer.tokenize(prompt))
        for normalize in (False, True):
            keys, maps = tc.compute_per_head_heat_maps(normalize=normalize)
            assert len(keys) == 25
            for (factor, layer, head), m in zip(keys, maps):
                ref = O.port_global_heat_map(store, 4096, n_tok, layer_idx=layer, head_idx=head, normalize=normalize)
                assert_elementwise(m, ref, 1e-3, 2e-4, f'{(factor, layer, head)} normalize={normalize}')


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16])
def test_latent96_pipeline_vs_oracle(dtype):
    """sample_size 96 (the 768-pixel models; latent_hw = 9216, daam/trace.py:32-33): layers at 9216 / 2304 / 576 query
    positions (partial 128-pixel tiles), (96, 96) global maps; oracle fed the identical Q/K."""
    pipe = make_pipeline(TINY96_SPEC, dtype=dtype, device=DEV, seed=5)
    prompt = 'a dog chasing a red ball on the beach'
    steps = 2
    with trace(pipe) as tc:
        assert tc.latent_hw == 9216
        rec = HookRecorder(tc)
        pipe(prompt, num_inference_steps=steps, generator=torch.Generator().manual_seed(13))
        store = rec.oracle_store()
        got = {k: v.clone() for k, v in tc.all_heat_maps}
        assert {k[0] for k in got} == {1, 2, 4} and {v.shape[-1] for v in got.values()} == {96, 48, 24}
        for key, ref in store:
            assert_elementwise(got[key], ref, RTOL[dtype], ATOL[dtype] * steps, f'key {key}')
        n_tok = len(pipe.tokenizer.tokenize(prompt))
        mult = 1 if dtype == torch.float32 else 10
        for kw in [{}, {'normalize': True}, {'factors': [4]}, {'layer_idx': 9, 'head_idx': 0}]:
            ref = O.port_global_heat_map(store, 9216, n_tok, **kw)
            out = tc.compute_global_heat_map(**kw).heat_maps
            assert out.shape == ref.shape == (n_tok + 2, 96, 96)
            assert_elementwise(out, ref, 1e-4 * mult, 1e-5 * steps * mult, f'{kw}')


def test_latent96_pipeline_vs_reference_fixture():
    """The verbatim reference's own run of the 96-latent pipeline (CPU fp32; tests/golden/pipeline_tiny96.npz) vs ours
    (GPU fp32). As for pipeline_tiny: Q/K differ by GPU-vs-CPU matmul rounding fed back through the UNet steps, hence the
    pipeline-level 1e-3 (of the map's max)."""
    from tests.util import rel_err
    fx = golden('pipeline_tiny96')
    pipe = make_pipeline(TINY96_SPEC, dtype=torch.float32, device=DEV, seed=int(fx['unet_seed']))
    with trace(pipe) as tc:
        pipe(str(fx['prompt']), num_inference_steps=int(fx['steps']),
             generator=torch.Generator().manual_seed(int(fx['gen_seed'])))
        keys = sorted(k for k, _ in tc.all_heat_maps)
        assert keys == sorted(tuple(k) for k in fx['keys'].tolist())
        sums = {k: float(v.double().sum()) for k, v in tc.all_heat_maps}
        for k, s in zip(fx['keys'].tolist(), fx['key_sums']):
            assert abs(sums[tuple(k)] - s) < 1e-4 * s
        g = tc.compute_global_heat_map()
        assert g.heat_maps.shape == (11, 96, 96)
        assert rel_err(g.heat_maps, fx['global']) < 1e-3
        assert rel_err(tc.compute_global_heat_map(normalize=True).heat_maps, fx['global_norm']) < 1e-3
        assert rel_err(tc.compute_global_heat_map(factors=[4]).heat_maps, fx['factors_4']) < 1e-3
        assert rel_err(g.compute_word_heat_map('ball').heatmap, fx['word_ball']) < 1e-3


def test_plan_cache_replays_and_invalidates():
    """daam_accumulate caches its launch plan by the verbatim daam_layer[] input: replays must keep accumulating, and a
    changed pointer / shape / flag must not hit a stale plan."""
    g = torch.Generator().manual_seed(3)
    hw, heads, d = 256, 2, 64
    qa, qb = [torch.randn(2, hw, heads * d, generator=g).half().to(DEV) for _ in range(2)]
    k = torch.randn(2, 77, heads * d, generator=g).half().to(DEV)
    acc = ops.new_accumulator(1, heads, hw, DEV)
    da, db = ops.make_layer_desc(qa, k, acc, heads, 0.125), ops.make_layer_desc(qb, k, acc, heads, 0.125)
    for desc in (da, da, db, da, db):
        ops.accumulate([desc], DEV)
    torch.cuda.synchronize()
    ref = 3 * oracle_layer_maps(qa, k, heads, 0.125) + 2 * oracle_layer_maps(qb, k, heads, 0.125)
    assert_elementwise(acc[0], ref, 1e-4, 5e-5, 'alternating plans')
    # more distinct inputs than the cache holds (32): eviction must not corrupt anything
    accs = [ops.new_accumulator(1, heads, hw, DEV) for _ in range(40)]
    for rep in range(2):
        for a in accs:
            ops.accumulate([ops.make_layer_desc(qa, k, a, heads, 0.125)], DEV)
    torch.cuda.synchronize()
    one = oracle_layer_maps(qa, k, heads, 0.125)
    for a in (accs[0], accs[17], accs[39]):
        assert_elementwise(a[0], 2 * one, 1e-4, 2e-5, 'evicted plans')


def test_early_loads_back_to_back_is_bit_identical_to_serialised_launches():
    """DAAM_ACC_EARLY_LOADS lets a launch's loads, MMAs and first softmax overlap the previous launch's tail; only the
    accumulator updates wait for it. The order of updates per element is unchanged, so 30 back-to-back launches on the
    same slabs must equal the serialised (no-PDL) sequence bit for bit, for both operand classes."""
    g = torch.Generator().manual_seed(17)
    for dtype in (torch.bfloat16, torch.float32):
        shapes = [(1024, 10), (256, 20), (4096, 5)]
        qs = [[torch.randn(2, hw, h * 64, generator=g).to(dtype).to(DEV) for hw, h in shapes] for _ in range(3)]
        ks = [[torch.randn(2, 77, h * 64, generator=g).to(dtype).to(DEV) for hw, h in shapes] for _ in range(3)]
        results = []
        for flags in (_native.ACC_NO_PDL, _native.ACC_EARLY_LOADS, _native.ACC_AUTO):
            accs = [ops.new_accumulator(1, h, hw, DEV) for hw, h in shapes]
            packs = [ops.pack([ops.make_layer_desc(q, k, a, h, 0.125) for q, k, a, (hw, h) in zip(qs[i], ks[i], accs, shapes)])
                     for i in range(3)]
            torch.cuda.synchronize()
            for step in range(30):
                ops.accumulate(packs[step % 3], DEV, flags=flags)
            torch.cuda.synchronize()
            results.append(accs)
        for a, b, c in zip(*results):
            assert torch.equal(a, b) and torch.equal(a, c)
        sums = results[1][0].double().sum(dim=(2, 3))
        assert torch.allclose(sums, torch.full_like(sums, 30.0 * shapes[0][0]), rtol=2e-5)


@pytest.mark.parametrize('dtype', [torch.float16, torch.float32])
def test_weighted_partition_of_mixed_head_dims(dtype):
    """One launch with 1-, 2-, 3- and 4-chunk layers of very different tile counts (the K-chunked instances partition by
    weight, some CTAs get no tile at all): every layer against the oracle, nothing outside the slabs touched."""
    g = torch.Generator().manual_seed(23)
    cases = [(16, 2, 256), (4096, 2, 40), (64, 3, 192), (1024, 1, 80), (256, 8, 160), (144, 2, 64), (16, 1, 8)]
    qs, ks, slabs, descs = [], [], [], []
    for hw, heads, d in cases:
        q = torch.randn(2, hw, heads * d, generator=g).to(dtype).to(DEV)
        k = torch.randn(2, 77, heads * d, generator=g).to(dtype).to(DEV)
        slab = torch.zeros(heads + 2, 77, hw, device=DEV)                # guard heads before and after
        qs.append(q), ks.append(k), slabs.append(slab)
        descs.append(ops.make_layer_desc(q, k, slab[1:-1].unsqueeze(0), heads, d ** -0.5))
    for rep in range(2):
        ops.accumulate(descs, DEV, flags=_native.ACC_FORCE_MMA | _native.ACC_EARLY_LOADS)
    torch.cuda.synchronize()
    for (hw, heads, d), q, k, slab in zip(cases, qs, ks, slabs):
        ref = 2 * oracle_layer_maps(q, k, heads, d ** -0.5)
        assert_elementwise(slab[1:-1], ref, RTOL[dtype], 2 * ATOL[dtype], f'hw{hw} H{heads} d{d}')
        assert float(slab[0].abs().max()) == 0.0 and float(slab[-1].abs().max()) == 0.0
