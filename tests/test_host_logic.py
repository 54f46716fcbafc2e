"""Host-side logic that needs no GPU: layer enumeration, hook plumbing, prompt/word bookkeeping, error behaviour."""
import pytest
import torch

import daam_b200
from daam_b200 import trace
from daam_b200.build import build
from daam_b200.hook import AggregateHooker, ObjectHooker, UNetCrossAttentionLocator
from daam_b200.ops import cond_half
from daam_b200.testing.synthetic import (SD21_SPEC, SDXL_SPEC, TINY_SPEC, SDPAProcessor, SyntheticUNet, WhitespaceTokenizer,
                                 make_pipeline)
from daam_b200.utils import compute_token_merge_indices


@pytest.fixture(scope='module', autouse=True)
def _built():
    build()


def meta_unet(spec):
    with torch.device('meta'):
        return SyntheticUNet(spec, body='skeleton')


def test_export_surface():
    for name in ['trace', 'set_seed', 'GlobalHeatMap', 'WordHeatMap', 'RawHeatMapCollection', 'ObjectHooker',
                 'AggregateHooker', 'UNetCrossAttentionLocator', 'compute_token_merge_indices', 'auto_device',
                 'auto_autocast', 'cache_dir', 'DiffusionHeatMapHooker']:
        assert hasattr(daam_b200, name), name
    assert daam_b200.trace is daam_b200.DiffusionHeatMapHooker


def test_locator_sd21_order_and_names():
    loc = UNetCrossAttentionLocator()
    layers = loc.locate(meta_unet(SD21_SPEC))
    assert len(layers) == 15
    assert loc.layer_names == ['up-attn-0', 'up-attn-1', 'up-attn-2'] * 3 + ['down-attn-0', 'down-attn-1'] * 3
    dims = [(l.to_q.in_features, l.heads) for l in layers]
    assert dims == [(1280, 20)] * 3 + [(640, 10)] * 3 + [(320, 5)] * 3 + [(320, 5)] * 2 + [(640, 10)] * 2 + [(1280, 20)] * 2
    loc_mid = UNetCrossAttentionLocator(locate_middle_block=True)
    assert len(loc_mid.locate(meta_unet(SD21_SPEC))) == 16 and loc_mid.layer_names[-1] == 'mid-attn-0'
    loc_low = UNetCrossAttentionLocator(restrict={0})
    assert len(loc_low.locate(meta_unet(SD21_SPEC))) == 6
    assert loc_low.layer_names == ['up-attn-0'] * 3 + ['down-attn-0'] * 3


def test_locator_sdxl_counts():
    unet = meta_unet(SDXL_SPEC)
    assert len(UNetCrossAttentionLocator().locate(unet)) == 60
    loc = UNetCrossAttentionLocator(locate_middle_block=True)
    layers = loc.locate(unet)
    assert len(layers) == 70
    assert [l.heads for l in layers[:30]] == [20] * 30 and [l.heads for l in layers[30:36]] == [10] * 6
    assert loc.layer_names[:3] == ['up-attn-0', 'up-attn-1', 'up-attn-2'] and loc.layer_names[29] == 'up-attn-29'


def test_object_hooker_patch_and_restore():
    class Target:
        def greet(self, x):
            return f'hi {x}'

    class Hooker(ObjectHooker):
        def _hook_impl(self):
            self.monkey_patch('greet', self._greet)
            self.monkey_patch('absent', self._greet, strict=False)

        def _greet(hk, target, x):
            return hk.monkey_super('greet', x).upper()

    t = Target()
    hk = Hooker(t)
    with pytest.raises(RuntimeError, match='Module is not hooked'):
        hk.unhook()
    with hk:
        assert t.greet('a') == 'HI A'
        with pytest.raises(RuntimeError, match='Already hooked module'):
            hk.hook()
    assert t.greet('a') == 'hi a' and 'greet' not in vars(t) or t.greet('a') == 'hi a'

    class Strict(ObjectHooker):
        def _hook_impl(self):
            self.monkey_patch('absent', lambda *_: None)

    with pytest.raises(AttributeError):
        Strict(t).hook()
    agg = AggregateHooker([Hooker(Target()), Hooker(Target())])
    with agg:
        assert all(h.hooked for h in agg.module)
    assert not any(h.hooked for h in agg.module)


def test_token_merge_indices():
    tok = WhitespaceTokenizer()
    assert compute_token_merge_indices(tok, 'A dog and a Dog', 'dog') == ([2, 5], None)
    assert compute_token_merge_indices(tok, 'a red ball', 'red ball') == ([2, 3], None)
    assert compute_token_merge_indices(tok, 'a red ball', 'x', word_idx=4) == ([5], 4)
    assert compute_token_merge_indices(tok, 'a red ball', 'ball', offset_idx=2) == ([5], None)
    with pytest.raises(ValueError, match='Search word zebra not found in prompt!'):
        compute_token_merge_indices(tok, 'a red ball', 'Zebra')


def test_cond_half_rule():
    assert cond_half(2, 5) == (1, 1, 0, 5)          # CFG pair: the conditional sample, all heads
    assert cond_half(16, 10) == (8, 8, 0, 10)       # batched prompts: [uncond x 8, cond x 8]
    assert cond_half(1, 8) == (0, 1, 4, 4)          # no guidance: the reference keeps the upper half of the heads
    with pytest.raises(RuntimeError):
        cond_half(3, 8)


def test_trace_installs_and_restores_processors():
    pipe = make_pipeline(TINY_SPEC)
    layers = UNetCrossAttentionLocator().locate(pipe.unet)
    check_inputs = pipe.check_inputs
    tc = trace(pipe)
    assert tc.latent_hw == 4096 and len(tc.layer_names) == 15
    with tc:
        assert all(l.processor is h for l, h in zip(layers, tc.module[:15]))
        assert [h.layer_idx for h in tc.module[:15]] == list(range(15))
        with pytest.raises(RuntimeError, match='No heat maps found. Did you forget to call'):
            tc.compute_global_heat_map()
        with pytest.raises(RuntimeError, match='No heat maps found for the given parameters.'):
            tc.compute_global_heat_map(layer_idx=3)
        with pytest.raises(ValueError, match='Only single prompt generation is supported'):
            pipe(['a cat', 'a dog'], num_inference_steps=1)
    assert all(isinstance(l.processor, SDPAProcessor) for l in layers)
    assert pipe.check_inputs == check_inputs
    tc.time_callback(0, 0, None)
    assert tc.time_idx == 1


def test_tracing_a_cpu_pipeline_fails_loudly():
    pipe = make_pipeline(TINY_SPEC)
    with trace(pipe, launch='layer') as tc:
        with pytest.raises(RuntimeError, match='no CPU fallback'):
            pipe('a cat', num_inference_steps=1)
        tc.all_heat_maps.clear()


def test_latent_hw_rule():
    for sample, expect in [(64, 4096), (128, 4096), (96, 9216)]:
        spec = TINY_SPEC.__class__(**{**TINY_SPEC.__dict__, 'sample_size': sample})
        pipe = make_pipeline(spec)
        assert trace(pipe).latent_hw == expect


def test_bad_options_are_rejected():
    pipe = make_pipeline(TINY_SPEC)
    with pytest.raises(ValueError):
        trace(pipe, launch='sometimes')


def test_generation_experiment_roundtrip_and_reference_dump(tmp_path):
    """Same folder layout as the reference (experiment.py:140-175); dumps pickled under the reference's class path load."""
    import sys
    import types
    from daam_b200 import GenerationExperiment
    tok = WhitespaceTokenizer()
    maps = torch.rand(5, 8, 8)
    exp = GenerationExperiment(image=None, global_heat_map=maps, prompt='a red ball', seed=7, id='p0', path=str(tmp_path),
                               tokenizer=tok).annotate('k', 1)
    exp.save()
    assert (tmp_path / 'p0' / 'generation.pt').exists() and (tmp_path / 'p0' / 'prompt.txt').read_text() == 'a red ball'
    assert GenerationExperiment.read_seed(tmp_path, 'p0') == 7 and GenerationExperiment.has_experiment(tmp_path, 'p0')
    back = GenerationExperiment.load(tmp_path / 'p0')
    assert back.prompt == 'a red ball' and torch.equal(back.global_heat_map, maps) and back.annotations == {'k': 1}
    assert back.heat_map().prompt == 'a red ball'
    # a dump written by the reference pickles the class as daam.experiment.GenerationExperiment
    fake = types.ModuleType('daam.experiment')
    fake.GenerationExperiment = type('GenerationExperiment', (), {'__module__': 'daam.experiment'})
    had_pkg, had = sys.modules.get('daam'), sys.modules.get('daam.experiment')
    if had_pkg is None:
        sys.modules['daam'] = types.ModuleType('daam')
    sys.modules['daam.experiment'] = fake
    try:
        obj = fake.GenerationExperiment()
        obj.__dict__.update(image=None, global_heat_map=maps, prompt='ref prompt', seed=1, id='.', path=None,
                            truth_masks=None, prediction_masks=None, annotations=None, subtype='.', tokenizer=None)
        (tmp_path / 'ref').mkdir()
        torch.save(obj, tmp_path / 'ref' / 'generation.pt')
    finally:
        if had is None:
            del sys.modules['daam.experiment']
        else:
            sys.modules['daam.experiment'] = had
        if had_pkg is None:
            del sys.modules['daam']
    ref = GenerationExperiment.load(tmp_path / 'ref')
    assert isinstance(ref, GenerationExperiment) and ref.prompt == 'ref prompt' and torch.equal(ref.global_heat_map, maps)


def test_bench_workload_definitions_match_the_survey():
    """13.80 Mpx and 135.1 MB of algorithmic traffic per SD-2.1 step (bf16), 15 / 60 traced layers."""
    import bench
    sd21, sdxl = bench.traced_layers('sd21'), bench.traced_layers('sdxl')
    assert len(sd21) == 15 and len(sdxl) == 60
    assert bench.px_per_step(sd21) == 13_798_400
    assert bench.literal_px_per_step(sd21) == 15 * 77 * 4096
    assert abs(bench.algorithmic_bytes_per_step(sd21, esize=2) / 1e6 - 135.05) < 0.01
    assert abs(bench.algorithmic_bytes_per_step(sd21, esize=4) / 1e6 - 159.71) < 0.01
    # SDXL default trace: 10 layers at 64^2 with 10 heads (30.57 MB each), 50 at 32^2 with 20 heads (15.43 MB each)
    assert abs(bench.algorithmic_bytes_per_step(sdxl, esize=2) / 1e6 - (10 * 30.57 + 50 * 15.43)) < 0.5
    sd15 = bench.traced_layers('sd15')
    assert len(sd15) == 15 and {h for _, h, _ in sd15} == {8} and {d for _, _, d in sd15} == {40, 80, 160}
    assert bench.px_per_step(sd21, n_prompts=8) == 8 * 13_798_400
    # BASELINE config 5, "all 70 cross-attn layers traced": SURVEY.md 8d gives 126.2 Mpx and 1 231.7 MB per step and prompt
    sdxl70 = bench.traced_layers('sdxl70')
    assert len(sdxl70) == 70 and sdxl70[:60] == sdxl
    assert abs(bench.px_per_step(sdxl70) / 1e6 - 126.2) < 0.1
    assert abs(bench.algorithmic_bytes_per_step(sdxl70, esize=2) / 1e6 - 1231.7) < 0.5


def test_bench_config_is_identical_for_both_arms_and_located_layers_match_the_workloads():
    """The `config` object both bench arms emit comes from one function; the layer lists the bench assumes are what the
    locator finds on the synthetic SDXL tree (60 by default, 70 with the mid block, reference: daam/hook.py:110-114)."""
    import argparse
    import bench
    from daam_b200.locate import UNetCrossAttentionLocator
    from daam_b200.testing.synthetic import SDXL_SPEC, SyntheticUNet
    args = argparse.Namespace(workload='sdxl70', prompts=2, dtype='fp16', steps=30)
    cfg = bench.shared_config(args, bench.traced_layers('sdxl70'), 8)
    assert cfg == bench.shared_config(args, bench.traced_layers('sdxl70'), 8)
    assert 'all 70 cross-attn layers' in cfg['workload'] and cfg['parallelism'].endswith('dp8')
    assert cfg['px_per_step'] == 2 * bench.px_per_step(bench.traced_layers('sdxl70'))
    with torch.device('meta'):
        unet = SyntheticUNet(SDXL_SPEC, body='skeleton')
    assert len(UNetCrossAttentionLocator().locate(unet)) == 60
    located = UNetCrossAttentionLocator(locate_middle_block=True).locate(unet)
    assert len(located) == 70
    shapes = [(m.heads, m.to_q.in_features) for m in located]
    # up blocks first (1280-wide x 30, 640-wide x 6), then down (640 x 4, 1280 x 20), the mid block last
    assert shapes == [(20, 1280)] * 30 + [(10, 640)] * 6 + [(10, 640)] * 4 + [(20, 1280)] * 20 + [(20, 1280)] * 10


def test_bench_clock_sampler_and_stdout_contract(capfd):
    """The pieces of bench.py the driver parses: throttle reasons / median clock from nvidia-smi rows, and exactly one
    JSON line on the real stdout even when libraries print to fd 1."""
    import json
    import os
    import time
    import bench
    s = bench.ClockSampler.__new__(bench.ClockSampler)
    s.proc = type('P', (), {'terminate': lambda self: None})()
    now = time.time()
    s.rows = [(now - 10, ['300', '1965', '80', '0', 'Not Active', 'Not Active', 'Not Active', 'Not Active']),
              (now - 1.0, ['1965', '1965', '900', '99', 'Not Active', 'Not Active', 'Not Active', 'Active']),
              (now - 0.9, ['1800', '1965', '950', '99', 'Not Active', 'Not Active', 'Not Active', 'Active']),
              (now - 0.8, ['1900', '1965', '950', '99', 'Not Active', 'Not Active', 'Not Active', 'Not Active'])]
    out = s.stop([(now - 2, now)])
    assert out == {'sm_mhz': 1900.0, 'sm_max_mhz': 1965.0, 'reasons': ['sw_power_cap'], 'samples': 3}
    bench._REAL_STDOUT = None
    bench.capture_stdout()
    try:
        os.write(1, b'NCCL version banner\\n')          # a library writing to fd 1 lands on stderr
        bench.emit({'metric': 'x', 'value': 1})
    finally:
        os.dup2(bench._REAL_STDOUT, 1)
        os.close(bench._REAL_STDOUT)
        bench._REAL_STDOUT = None
    captured = capfd.readouterr()
    assert captured.out.strip().splitlines() == [json.dumps({'metric': 'x', 'value': 1})]
    assert 'NCCL version banner' in captured.err
