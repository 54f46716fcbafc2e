"""Parity of the finalize / word-map / expand kernels with the reference's outputs (golden fixtures) and the oracle.

Tolerance: rtol 1e-5 of the output's max (+1e-6 abs): only fp32 summation order differs (SURVEY.md section 8c proposes
rtol 1e-4 for the global map; the kernels are well inside it)."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from daam_b200 import _native
from daam_b200.heatmap import GlobalHeatMap
from daam_b200.testing.synthetic import WhitespaceTokenizer
from oracle import daam_oracle as O
from tests.util import golden, rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def fixture_groups(fx, select=lambda f, l, h: True, head_sel=None):
    """Key tensors of the fixture grouped per layer (heads stacked): what the tracer's slabs look like."""
    by_layer = {}
    for name in fx.files:
        if name.startswith('key_'):
            f, l, h = (int(v) for v in name.split('_')[1:])
            by_layer.setdefault(l, {})[h] = (f, torch.from_numpy(fx[name]))
    groups, keep = [], []
    for l in sorted(by_layer):
        heads = by_layer[l]
        f = heads[0][0]
        if not select(f, l, None):
            continue
        if head_sel is not None and head_sel not in heads:
            continue
        t = torch.stack([heads[h][1] for h in sorted(heads)]).to(DEV)       # [H, tokens, h, w]
        keep.append(t)
        groups.append(_native.DaamKeyGroup(acc=t.data_ptr(), heads=t.shape[0], h=t.shape[2], w=t.shape[3],
                                           tokens=t.shape[1], head_sel=-1 if head_sel is None else head_sel,
                                           reserved=0))
    return groups, keep


def run_finalize(groups, n_rows, normalize=False, x=64):
    out = torch.empty(n_rows, x, x, device=DEV)
    _native.finalize(groups, x, n_rows, normalize, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize('case,select,head_sel,normalize', [
    ('global', lambda f, l, h: True, None, False),
    ('global_norm', lambda f, l, h: True, None, True),
    ('factors_2_4', lambda f, l, h: f in (2, 4), None, False),
    ('layer_1', lambda f, l, h: l == 1, None, False),
    ('head_1', lambda f, l, h: True, 1, False),
    ('layer_2_head_0', lambda f, l, h: l == 2, 0, False),
])
def test_finalize_golden(case, select, head_sel, normalize):
    fx = golden('finalize')
    groups, keep = fixture_groups(fx, select, head_sel)
    ref = fx[case]
    out = run_finalize(groups, ref.shape[0], normalize)
    assert rel_err(out, ref) < 1e-5, case
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=1e-4, atol=1e-6)


def test_factor_one_is_identity_and_clamped():
    """bicubic at scale 1 is the identity: a lone 64x64 key comes back unchanged, negatives clamped."""
    g = torch.Generator().manual_seed(1)
    key = torch.randn(1, 77, 64, 64, generator=g).to(DEV)
    grp = [_native.DaamKeyGroup(acc=key.data_ptr(), heads=1, h=64, w=64, tokens=77, head_sel=-1, reserved=0)]
    out = run_finalize(grp, 77)
    assert torch.equal(out, key[0].clamp(min=0))


@pytest.mark.parametrize('h,x', [(32, 64), (16, 64), (8, 64), (48, 96), (24, 96), (32, 32)])
def test_bicubic_matches_float64_math(h, x):
    g = torch.Generator().manual_seed(h + x)
    key = torch.exp(2.0 * torch.randn(2, 77, h, h, generator=g))
    kd = key.to(DEV)
    grp = [_native.DaamKeyGroup(acc=kd.data_ptr(), heads=2, h=h, w=h, tokens=77, head_sel=-1, reserved=0)]
    out = run_finalize(grp, 20, x=x)
    ref = O.math_global_heat_map([key[0].numpy(), key[1].numpy()], x, 20)
    assert rel_err(out, ref) < 1e-5


def test_finalize_rejects_empty_selection():
    out = torch.empty(4, 64, 64, device=DEV)
    with pytest.raises(_native.NativeError) as e:
        _native.finalize([], 64, 4, False, out.data_ptr(), 0)
    assert e.value.code == _native.E_INVALID


def test_word_heat_map_and_expand_golden():
    fx = golden('finalize')
    tok = WhitespaceTokenizer()
    g = torch.from_numpy(fx['global']).to(DEV)
    ghm = GlobalHeatMap(tok, str(fx['prompt']), g)
    w = ghm.compute_word_heat_map('three')
    assert w.word == 'three' and w.heatmap.is_cuda
    assert rel_err(w.heatmap, fx['word_three']) < 1e-6
    multi = GlobalHeatMap(tok, 'red ball and red car', g).compute_word_heat_map('red')
    assert rel_err(multi.value, fx['word_red_multi']) < 1e-6
    by_idx = ghm.compute_word_heat_map('anything', word_idx=2)           # row 3 == 'three'
    assert torch.equal(by_idx.heatmap, w.heatmap)
    with pytest.raises(ValueError, match='Search word zebra not found in prompt!'):
        ghm.compute_word_heat_map('zebra')
    with pytest.raises(IndexError):
        ghm.compute_word_heat_map('x', word_idx=40)
    img = SimpleNamespace(size=(96, 80))
    for case, kw, tol in [('expand', {}, 1e-5), ('expand_abs', {'absolute': True}, 1e-5)]:
        got = w.expand_as(img, **kw)
        assert not got.is_cuda and got.shape == (96, 80)
        assert rel_err(got, fx[case]) < tol, case
    thr = w.expand_as(img, threshold=0.4)
    mism = (thr.numpy() != fx['expand_thr']).mean()
    assert mism < 1e-3      # binarisation may flip pixels that sit within rounding of the threshold
    assert w.compute_ioa(w) == pytest.approx(float((w.heatmap ** 2).sum() / w.heatmap.sum()), rel=1e-5)


@pytest.mark.parametrize('x,sides', [(64, [64, 32, 16]), (96, [96, 48, 24]), (64, [16]), (32, [32, 8])])
def test_fast_and_generic_finalize_agree(monkeypatch, x, sides):
    """The phase-structured kernel for integer factors 1/2/4 vs the generic gather kernel (DAAM_FINALIZE_GENERIC=1),
    on SD-like key sets (several heads per resolution, head filter, normalisation, partial 96-latent bands)."""
    g = torch.Generator().manual_seed(x)
    keep, groups = [], []
    for i, side in enumerate(sides):
        heads = 3 + i
        t = torch.exp(2.0 * torch.randn(heads, 77, side, side, generator=g)).to(DEV)
        keep.append(t)
        groups.append(_native.DaamKeyGroup(acc=t.data_ptr(), heads=heads, h=side, w=side, tokens=77, head_sel=-1,
                                           reserved=0))
    for head_sel in (-1, 1):
        for grp in groups:
            grp.head_sel = head_sel
        for normalize in (False, True):
            monkeypatch.setenv('DAAM_FINALIZE_GENERIC', '0')
            fast = run_finalize(groups, 23, normalize, x=x)
            monkeypatch.setenv('DAAM_FINALIZE_GENERIC', '1')
            slow = run_finalize(groups, 23, normalize, x=x)
            assert rel_err(fast, slow) < 2e-6
            keys = [k[h].cpu().numpy() for k in keep for h in (range(k.shape[0]) if head_sel < 0 else [head_sel])]
            ref = O.math_global_heat_map(keys, x, 23, normalize)
            assert rel_err(fast, ref) < 1e-5


@pytest.mark.parametrize('n_rows', [1, 2, 3])
def test_short_prompts_and_degenerate_normalisation(n_rows):
    """rows[:n] with n = 1, 2: `maps[1:-1]` is empty, the reference divides by 1e-6 (trace.py:129-130); n = 3: one row."""
    g = torch.Generator().manual_seed(n_rows)
    key = torch.rand(2, 77, 32, 32, generator=g)
    kd = key.to(DEV)
    grp = [_native.DaamKeyGroup(acc=kd.data_ptr(), heads=2, h=32, w=32, tokens=77, head_sel=-1, reserved=0)]
    store = [((2, 0, h), key[h]) for h in range(2)]
    for normalize in (False, True):
        out = run_finalize(grp, n_rows, normalize)
        ref = O.port_global_heat_map(store, 4096, n_rows - 2, normalize=normalize)
        assert out.shape == ref.shape == (n_rows, 64, 64)
        assert rel_err(out, ref) < 1e-5


def test_iou_ioa_match_the_reference_formulas():
    """daam/evaluate.py:14-35, including the resize-and-binarise branch when the masks differ in size."""
    import torch.nn.functional as F
    from daam_b200 import compute_ioa, compute_iou
    g = torch.Generator().manual_seed(5)
    a = (torch.rand(64, 64, generator=g) > 0.5).float()
    b = (torch.rand(64, 64, generator=g) > 0.5).float()
    inter = (a * b).sum()
    assert compute_iou(a.to(DEV), b.to(DEV)) == pytest.approx((inter / (a.sum() + b.sum() - inter + 1e-8)).item(), rel=1e-6)
    assert compute_ioa(a.to(DEV), b.to(DEV)) == pytest.approx((inter / (a.sum() + 1e-8)).item(), rel=1e-6)
    # resize-and-binarise branch. A pixel whose upsampled value sits within rounding of the threshold may flip between
    # the CPU and the GPU interpolation, so the inputs are chosen (by seed, checked here) to keep every upsampled value
    # at least 1e-4 away from it: then the masks, and therefore the ratios, must agree to rounding of the sums.
    for seed in range(5, 50):
        g = torch.Generator().manual_seed(seed)
        small = torch.rand(16, 16, generator=g) * 2.0
        big = (torch.rand(64, 64, generator=g) > 0.5).float()
        up_f = F.interpolate(small[None, None], size=(64, 64), mode='bicubic').squeeze()
        if float((up_f - 1).abs().min()) > 1e-4:
            break
    else:
        raise AssertionError('no seed keeps the upsampled map away from the threshold')
    up = (up_f >= 1).float()
    i2 = (up * big).sum()
    assert compute_iou(small.to(DEV), big.to(DEV)) == pytest.approx((i2 / (up.sum() + big.sum() - i2 + 1e-8)).item(), rel=1e-6)
    assert compute_ioa(small.to(DEV), big.to(DEV)) == pytest.approx((i2 / (up.sum() + 1e-8)).item(), rel=1e-6)


def test_expand_words_fused_matches_reference_fixture_and_the_per_word_loop():
    """daam_expand_words (one launch for a word list: row gather-mean -> bicubic to the image size -> min/max ->
    normalise / threshold; daam/heatmap.py:121-123 + 77-93) against the verbatim reference's expand_as outputs and
    against this package's own per-word compute_word_heat_map().expand_as() loop."""
    fx = golden('finalize')
    tok = WhitespaceTokenizer()
    g = torch.from_numpy(fx['global']).to(DEV)
    prompt = str(fx['prompt'])
    ghm = GlobalHeatMap(tok, prompt, g)
    img = SimpleNamespace(size=(96, 80))
    words = prompt.split()
    before = _native.launch_count()
    whms, exp = ghm.expand_words(words, img)
    assert _native.launch_count() - before == 1                       # one launch for the whole list
    assert not exp.is_cuda and exp.shape == (len(words), 96, 80)
    i3 = words.index('three')
    assert rel_err(whms[i3].heatmap, fx['word_three']) < 1e-6 and whms[i3].word == 'three'
    assert rel_err(exp[i3], fx['expand']) < 1e-5
    _, exp_abs = ghm.expand_words(['three'], img, absolute=True)
    assert rel_err(exp_abs[0], fx['expand_abs']) < 1e-5
    _, exp_thr = ghm.expand_words(['three'], img, threshold=0.4)
    assert (exp_thr[0].numpy() != fx['expand_thr']).mean() < 1e-3
    for i, w in enumerate(words):                                       # == the per-word loop, bit for bit
        single = ghm.compute_word_heat_map(w)
        assert torch.equal(single.heatmap, whms[i].heatmap), w
        assert torch.equal(single.expand_as(img), exp[i]), w
    # multi-row words (two occurrences), explicit word indices, device-resident result, image-sized output
    multi = GlobalHeatMap(tok, 'red ball and red car', g)
    big = SimpleNamespace(size=(512, 512))
    whm2, dev_out = multi.expand_words(['red', 'car', 'x'], big, word_idx=[None, None, 2], to_cpu=False)
    assert dev_out.is_cuda and dev_out.shape == (3, 512, 512)
    assert rel_err(whm2[0].heatmap, fx['word_red_multi']) < 1e-6
    for i, (w, idx) in enumerate([('red', None), ('car', None), ('x', 2)]):
        ref = multi.compute_word_heat_map(w, word_idx=idx).expand_as(big)
        assert torch.equal(ref, dev_out[i].cpu()), w
        assert float(dev_out[i].min()) == 0.0 and abs(float(dev_out[i].max()) - 1.0) < 1e-6
    with pytest.raises(ValueError, match='Search word zebra not found in prompt!'):
        ghm.expand_words(['one', 'zebra'], img)
    assert ghm.expand_words([], img)[0] == []


def test_expand_words_many_words_and_oracle():
    """More (word, chunk) CTAs than one cooperative grid holds at the largest chunk count: the host batches; every word
    against the oracle's port of compute_word_heat_map + expand_as."""
    g = torch.Generator().manual_seed(4)
    n_rows = 77
    maps = torch.exp(torch.randn(n_rows, 64, 64, generator=g))
    prompt = ' '.join(f'w{i}' for i in range(75))
    tok = WhitespaceTokenizer()
    ghm = GlobalHeatMap(tok, prompt, maps.to(DEV))
    img = SimpleNamespace(size=(128, 96))
    words = prompt.split()
    whms, exp = ghm.expand_words(words, img)
    assert exp.shape == (75, 128, 96)
    for i in (0, 1, 37, 74):
        ref_w = O.port_word_heat_map(maps, tok, prompt, words[i])
        assert rel_err(whms[i].heatmap, ref_w) < 1e-6
        assert rel_err(exp[i], O.port_expand_as(ref_w, (128, 96))) < 1e-5
