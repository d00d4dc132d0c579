"""every parameter of the drop-in modules is drawn from the reference's initial
distribution — checked statistically (moments / support against the formula the cited reference line implies) and,
where the unmodified reference can be imported (build container), against the moments of the reference's own
freshly constructed modules.  CPU only: constructors do not touch the device."""
import math

import pytest
import torch


def _moments(t):
    t = t.detach().double().flatten()
    return float(t.mean()), float(t.std()), float(t.min()), float(t.max())


def _check_normal(t, std, what):
    m, s, _, _ = _moments(t)
    n = t.numel()
    assert abs(m) < 5 * std / math.sqrt(n), (what, "mean", m)
    assert abs(s / std - 1) < 5 / math.sqrt(2 * n) + 1e-3, (what, "std", s, std)
    # a normal sample this large exceeds 3 sigma somewhere: rules out a uniform of the same variance
    if n >= 20000:
        assert float(t.detach().abs().max()) > 3 * std, what


def _check_uniform(t, bound, what):
    m, s, lo, hi = _moments(t)
    n = t.numel()
    assert lo >= -bound and hi <= bound, (what, lo, hi, bound)
    assert abs(s / (bound / math.sqrt(3)) - 1) < 5 / math.sqrt(n) + 2e-3, (what, "std", s)
    assert abs(m) < 5 * bound / math.sqrt(3 * n), (what, "mean", m)
    assert hi > 0.98 * bound and lo < -0.98 * bound, (what, "support not reached")


def test_initial_distributions_follow_the_reference_formulas():
    import gripnet_b200 as gb
    torch.manual_seed(3)
    gcn = gb.myGCN(300, 200)
    _check_uniform(gcn.weight, math.sqrt(6.0 / 500), "myGCN.weight glorot (layers.py:42-44)")
    assert float(gcn.bias.abs().sum()) == 0.0                                   # layers.py:46-47
    for after_relu in (False, True):
        r = gb.myRGCN(120, 90, 40, 32, after_relu=after_relu)
        _check_normal(r.att, 1.0 / math.sqrt(32), "myRGCN.att (layers.py:152)")
        std = 2.0 / 120 if after_relu else 1.0 / math.sqrt(120)                 # layers.py:154-160
        _check_normal(r.basis, std, f"myRGCN.basis after_relu={after_relu}")
        _check_normal(r.root, std, f"myRGCN.root after_relu={after_relu}")
        assert r.bias is None                                                   # bias=False by default, :128
    h = gb.homoGraph([64, 32, 16], start_graph=True, in_dim=2000)
    _check_normal(h.embedding, 1.0, "homoGraph.embedding (layers.py:249-250)")
    hr = gb.homoGraph([48, 32, 32], multi_relational=True, n_rela=16, n_base=8)
    assert [c.after_relu for c in hr.conv_list] == [False, True]                # layers.py:232
    ig = gb.interGraph(64, 16, 3000, target_feat_dim=32)
    _check_normal(ig.target_feat, 1.0, "interGraph.target_feat (layers.py:359-360)")
    _check_normal(ig.target_feat_down, 1.0, "interGraph.target_feat_down (layers.py:348-353)")
    _check_uniform(ig.conv.weight, math.sqrt(6.0 / 80), "interGraph.conv.weight")
    dm = gb.multiRelaInnerProductDecoder(80, 400)
    _check_normal(dm.weight, 1.0 / math.sqrt(80), "DistMult weight (decoder.py:25-26)")
    mc = gb.multiClassInnerProductDecoder(288, 100)
    _check_uniform(mc.weight, math.sqrt(6.0 / 388), "multiclass weight glorot (decoder.py:47-49)")
    enc = gb.encoder.RGCN(500, 64, 32, 16, 8, 4)
    _check_normal(enc.embedding, 1.0, "encoder.RGCN.embedding (encoder.py:11-12)")
    assert (enc.rgcn1.after_relu, enc.rgcn2.after_relu) == (False, True)        # encoder.py:14-18


def test_initial_distributions_match_the_reference_modules():
    """Same constructor arguments on both sides: every parameter has the same shape and (to sampling error)
    the same mean / standard deviation / support as the reference's own initialiser."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference checkout not present (GPU box)")
    import gripnet_b200 as gb
    L, D, _ = ref_loader.load()
    torch.manual_seed(11)
    pairs = [
        (gb.myGCN(300, 200), L.myGCN(300, 200)),
        (gb.myRGCN(120, 90, 40, 32, after_relu=False), L.myRGCN(120, 90, 40, 32, after_relu=False)),
        (gb.myRGCN(120, 90, 40, 32, after_relu=True), L.myRGCN(120, 90, 40, 32, after_relu=True)),
        (gb.homoGraph([64, 32, 16], start_graph=True, in_dim=2000), L.homoGraph([64, 32, 16], start_graph=True, in_dim=2000)),
        (gb.homoGraph([48, 32, 32], multi_relational=True, n_rela=16, n_base=8),
         L.homoGraph([48, 32, 32], multi_relational=True, n_rela=16, n_base=8)),
        (gb.interGraph(64, 16, 3000, target_feat_dim=32), L.interGraph(64, 16, 3000, target_feat_dim=32)),
        (gb.multiRelaInnerProductDecoder(80, 400), D.multiRelaInnerProductDecoder(80, 400)),
        (gb.multiClassInnerProductDecoder(288, 100), D.multiClassInnerProductDecoder(288, 100)),
    ]
    for ours, ref in pairs:
        po, pr = dict(ours.named_parameters()), dict(ref.named_parameters())
        assert sorted(po) == sorted(pr), type(ours).__name__
        for k in po:
            assert po[k].shape == pr[k].shape, (type(ours).__name__, k)
            mo, so, lo, ho = _moments(po[k])
            mr, sr, lr, hr = _moments(pr[k])
            n = po[k].numel()
            if sr == 0.0:                                     # zero-initialised biases
                assert so == 0.0, (type(ours).__name__, k)
                continue
            assert abs(so / sr - 1) < 8 / math.sqrt(n) + 5e-3, (type(ours).__name__, k, so, sr)
            assert abs(mo - mr) < 8 * sr / math.sqrt(n), (type(ours).__name__, k, mo, mr)
            # same family: kurtosis 1.8 for a uniform, 3 for a normal
            if n >= 20000:
                def kurt(t):
                    t = t.detach().double().flatten()
                    c = t - t.mean()
                    return float((c ** 4).mean() / (c ** 2).mean() ** 2)
                assert abs(kurt(po[k]) - kurt(pr[k])) < 0.25, (type(ours).__name__, k, kurt(po[k]), kurt(pr[k]))
