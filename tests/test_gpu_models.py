"""GPU: the drop-in modules against the reference's golden fixtures and the oracle port.

Golden fixtures = outputs of the unmodified reference (tests/golden/make_golden.py).
Tolerance 1e-5 relative (max|a-b|/max|b| per tensor) for embeddings, scores, losses
and every parameter gradient (north_star).
"""
import numpy as np
import pytest
import torch

from golden_util import load_case, load_grouped, rel_err
from oracle import port, synth

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _dev():
    return torch.device("cuda:0")


def _check_grads(model, golden, tol=TOL):
    named = dict(model.named_parameters())
    for k, g in golden.items():
        assert named[k].grad is not None, k
        err = rel_err(named[k].grad, g)
        assert err < tol, (k, err)


def _pose_model(c, **kw):
    from gripnet_b200.pipelines import PoseModel, load_flat_params, to_device
    g = c["in"]
    m = load_flat_params(PoseModel(g["n_g"], g["n_d"], g["n_rel"], **kw), c["p"]).to(_dev())
    return m, to_device(g, _dev())


@pytest.mark.parametrize("name", ["pose_small", "pose_small_weighted"])
def test_pose_matches_reference_golden(name):
    c = load_case(name)
    m, data = _pose_model(c)
    loss, z, pos, neg = m(data)
    loss.backward()
    assert rel_err(z, c["out"]["z"]) < TOL
    assert rel_err(pos, c["out"]["pos"]) < TOL and rel_err(neg, c["out"]["neg"]) < TOL
    assert rel_err(loss, c["out"]["loss"]) < TOL
    _check_grads(m, c["grad"])


def test_pose_stagewise_and_torch_loss_expression():
    """Stage outputs, and the script's own torch loss expression (GripNet-pose.py:140-142) on our scores."""
    c = load_case("pose_small")
    m, data = _pose_model(c)
    z_gg = m.gg(None, data["gg_edge_index"], if_catout=True)
    assert rel_err(z_gg, c["out"]["z_gg"]) < TOL
    z_gd = m.gd(z_gg, data["gd_edge_index"], mod="cat", if_relu=True)
    assert rel_err(z_gd, c["out"]["z_gd"]) < TOL
    z = m.dd(z_gd, data["dd_edge_index"], edge_type=data["dd_edge_type"], range_list=data["dd_range_list"],
             if_catout=True)
    pos = m.dmt(z, data["dd_edge_index"], data["dd_edge_type"])
    neg = m.dmt(z, data["neg_edge_index"], data["dd_edge_type"])
    loss = -torch.log(pos + 1e-13).mean() - torch.log(1 - neg + 1e-13).mean()
    loss.backward()
    assert rel_err(loss, c["out"]["loss"]) < TOL
    _check_grads(m, c["grad"])


def test_aminer_matches_reference_golden():
    from gripnet_b200.pipelines import AminerModel, load_flat_params, to_device
    c = load_case("aminer_small")
    g = c["in"]
    m = load_flat_params(AminerModel(g["n_p"], g["n_a"], g["n_class"], pp=(32, 16, 16), pa=(16, 16),
                                     aa_hid=(32, 8)), c["p"]).to(_dev())
    loss, z, score = m(to_device(g, _dev()))
    loss.backward()
    assert rel_err(z, c["out"]["z"]) < TOL and rel_err(score, c["out"]["score"]) < TOL
    assert rel_err(loss, c["out"]["loss"]) < TOL
    _check_grads(m, c["grad"])


def test_freebase_d_matches_reference_golden():
    from gripnet_b200.pipelines import FreebaseDModel, load_flat_params, to_device
    c = load_case("freebase_d_small")
    g = c["in"]
    m = load_flat_params(FreebaseDModel(g["n_p"], g["n_q"], g["n_a"], g["n_class"], pp=(32, 16, 16), pa=(16, 16),
                                        aa_out=8), c["p"]).to(_dev())
    loss, z, score = m(to_device(g, _dev()))
    loss.backward()
    assert rel_err(z, c["out"]["z"]) < TOL and rel_err(score, c["out"]["score"]) < TOL
    assert rel_err(loss, c["out"]["loss"]) < TOL
    _check_grads(m, c["grad"])


# ----------------------------------------------------------------- module variants
def _split(c):
    p = {k[2:]: v for k, v in c.items() if k.startswith("p.")}
    g = {k[5:]: v for k, v in c.items() if k.startswith("grad.")}
    return p, g


def _w(out):
    return torch.linspace(-1, 1, out.numel(), device=out.device).view_as(out)


@pytest.mark.parametrize("tag,tdim,tfd,mod,relu", [("add_eq", 16, 16, "add", True), ("add_down", 16, 12, "add", True),
                                                   ("cat_norelu", 16, 8, "cat", False)])
def test_inter_variants(tag, tdim, tfd, mod, relu):
    import gripnet_b200 as gb
    c = load_grouped("variants")["inter_" + tag]
    p, g = _split(c)
    m = gb.interGraph(24, tdim, p["target_feat"].shape[0], target_feat_dim=tfd)
    m.load_state_dict(p)
    m = m.to(_dev())
    x = c["x"].to(_dev()).requires_grad_(True)
    ew = c["edge_weight"].to(_dev()) if "edge_weight" in c else None
    out = m(x, c["edge_index"].to(_dev()), edge_weight=ew, if_relu=relu, mod=mod)
    (out * _w(out)).sum().backward()
    assert rel_err(out, c["out"]) < TOL and rel_err(x.grad, c["grad_x"]) < TOL
    _check_grads(m, g)


def test_homo_nocat_weighted():
    import gripnet_b200 as gb
    c = load_grouped("variants")["homo_nocat"]
    p, g = _split(c)
    m = gb.homoGraph([20, 16, 8])
    m.load_state_dict(p)
    m = m.to(_dev())
    x = c["x"].to(_dev()).requires_grad_(True)
    out = m(x, c["edge_index"].to(_dev()), edge_weight=c["edge_weight"].to(_dev()), if_catout=False)
    (out * _w(out)).sum().backward()
    assert rel_err(out, c["out"]) < TOL and rel_err(x.grad, c["grad_x"]) < TOL
    _check_grads(m, g)


def test_rgcn_variants():
    import gripnet_b200 as gb
    v = load_grouped("variants")
    c = v["rgcn2"]
    p, g = _split(c)
    m = gb.homoGraph([12, 20, 8], multi_relational=True, n_rela=4, n_base=6)
    m.load_state_dict(p)
    m = m.to(_dev())
    x = c["x"].to(_dev()).requires_grad_(True)
    out = m(x, c["edge_index"].to(_dev()), edge_type=c["edge_type"].to(_dev()),
            range_list=c["range_list"].to(_dev()), if_catout=True)
    (out * _w(out)).sum().backward()
    assert rel_err(out, c["out"]) < TOL and rel_err(x.grad, c["grad_x"]) < TOL
    _check_grads(m, g)
    c = v["rgcn_bias"]
    p, g = _split(c)
    conv = gb.myRGCN(12, 10, 4, 6, after_relu=False, bias=True)
    conv.load_state_dict(p)
    conv = conv.to(_dev())
    x = c["x"].to(_dev()).requires_grad_(True)
    out = conv(x, c["edge_index"].to(_dev()), c["edge_type"].to(_dev()), c["range_list"])   # CPU range_list is fine
    (out * _w(out)).sum().backward()
    assert rel_err(out, c["out"]) < TOL and rel_err(x.grad, c["grad_x"]) < TOL
    _check_grads(conv, g)


def test_rgcn_prologue_is_bit_identical_and_ignores_stale_parameters():
    """``homoGraph.prologue`` (W[r] and the tensor-core operand image built ahead, on a background stream) changes
    where the parameter-only work runs, not its result: outputs and gradients are bit-identical with and without it;
    a prologue taken BEFORE an in-place parameter update is discarded (version check), not used."""
    import gripnet_b200 as gb
    from gripnet_b200 import ops
    v = load_grouped("variants")
    c = v["rgcn2"]
    p, _ = _split(c)
    m = gb.homoGraph([12, 20, 8], multi_relational=True, n_rela=4, n_base=6)
    m.load_state_dict(p)
    m = m.to(_dev())
    args = (c["edge_index"].to(_dev()),)
    kw = dict(edge_type=c["edge_type"].to(_dev()), range_list=c["range_list"].to(_dev()), if_catout=True)

    def run(prologue):
        for q in m.parameters():
            q.grad = None
        x = c["x"].to(_dev()).requires_grad_(True)
        if prologue:
            assert m.prologue(x.size(0)) is not None
            assert len(ops.RelPrologue._pending) == 2           # one entry per relational layer
        out = m(x, *args, **kw)
        assert len(ops.RelPrologue._pending) == 0               # taken by the forward
        (out * _w(out)).sum().backward()
        torch.cuda.synchronize()
        return [out.detach().clone(), x.grad.clone()] + [q.grad.clone() for q in m.parameters()]

    plain, ahead = run(False), run(True)
    for a, b in zip(plain, ahead):
        assert torch.equal(a, b)
    assert rel_err(ahead[0], c["out"]) < TOL
    # stale prologue: parameters rewritten in place after it was taken
    m.prologue(c["x"].size(0))
    with torch.no_grad():
        for conv in m.conv_list:
            conv.att.mul_(2.0)
    x = c["x"].to(_dev())
    stale = m(x, *args, **kw)
    fresh = m(x, *args, **kw)
    assert torch.equal(stale, fresh) and not torch.equal(stale.detach(), ahead[0])


def test_gcn_improved_uncached_and_norm_api():
    import gripnet_b200 as gb
    c = load_grouped("variants")["gcn_improved"]
    p, g = _split(c)
    conv = gb.myGCN(10, 6, improved=True, cached=False, bias=False)
    conv.load_state_dict(p)
    conv = conv.to(_dev())
    x = c["x"].to(_dev()).requires_grad_(True)
    ei = c["edge_index"].to(_dev())
    out = conv(x, ei)
    (out * _w(out)).sum().backward()
    assert rel_err(out, c["out"]) < TOL and rel_err(x.grad, c["grad_x"]) < TOL
    _check_grads(conv, g)
    ei2, nrm = gb.myGCN.norm(ei, 40, None, improved=True)
    ei_ref, nrm_ref = port.gcn_norm(c["edge_index"], 40, None, improved=True)
    assert torch.equal(ei2.cpu(), ei_ref) and rel_err(nrm, nrm_ref) < 1e-6
    cached = conv.cached_result
    assert torch.equal(cached[0].cpu(), ei_ref)


def test_decoder_variants():
    import gripnet_b200 as gb
    v = load_grouped("variants")
    c = v["dmt_raw"]
    dec = gb.multiRelaInnerProductDecoder(20, 3)
    dec.load_state_dict({"weight": c["p.weight"]})
    dec = dec.to(_dev())
    z = c["z"].to(_dev()).requires_grad_(True)
    out = dec(z, c["edge_index"].to(_dev()), c["edge_type"].to(_dev()), sigmoid=False)
    (out * _w(out)).sum().backward()
    assert rel_err(out, c["out"]) < TOL and rel_err(z.grad, c["grad_z"]) < TOL
    assert rel_err(dec.weight.grad, c["grad.weight"]) < TOL
    for tag, sm in (("mcip_raw", False), ("mcip_softmax", True)):
        c = v[tag]
        dec = gb.multiClassInnerProductDecoder(20, 7)
        dec.load_state_dict({"weight": c["p.weight"]})
        dec = dec.to(_dev())
        z = c["z"].to(_dev()).requires_grad_(True)
        out = dec(z, c["node_list"].to(_dev()), softmax=sm)
        (out * _w(out)).sum().backward()
        assert rel_err(out, c["out"]) < TOL and rel_err(z.grad, c["grad_z"]) < TOL
        assert rel_err(dec.weight.grad, c["grad.weight"]) < TOL


# ----------------------------------------------------------------- larger, vs the oracle port
def test_pose_medium_matches_port_and_is_deterministic():
    from gripnet_b200.pipelines import PoseModel, load_flat_params, to_device
    g = synth.pose_medium()
    p = synth.pose_params(g)
    pl = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    loss_ref, z_ref, pos_ref, neg_ref = port.pose_forward(pl, g)
    loss_ref.backward()
    m = load_flat_params(PoseModel(g["n_g"], g["n_d"], g["n_rel"]), p).to(_dev())
    data = to_device(g, _dev())
    runs = []
    for _ in range(2):
        m.zero_grad(set_to_none=True)
        loss, z, pos, neg = m(data)
        loss.backward()
        runs.append((loss.detach().clone(), z.detach().clone(),
                     {k: v.grad.clone() for k, v in m.named_parameters() if v.grad is not None}))
    assert rel_err(runs[0][1], z_ref) < TOL and rel_err(runs[0][0], loss_ref) < TOL
    # target_feat_down is unused with mod="cat": no gradient in the reference either
    assert set(runs[0][2]) == {k for k, v in pl.items() if v.grad is not None}
    for k, v in runs[0][2].items():
        err = rel_err(v, pl[k].grad)
        assert err < 2e-5, (k, err)       # hub rows (degree ~1e3) reorder fp32 sums: SURVEY §7 "tolerance"
    # bit-identical second run: no data atomics anywhere in fwd or bwd
    assert torch.equal(runs[0][0], runs[1][0]) and torch.equal(runs[0][1], runs[1][1])
    for k in runs[0][2]:
        assert torch.equal(runs[0][2][k], runs[1][2][k]), k


def test_api_contract():
    import gripnet_b200 as gb
    d = _dev()
    m = gb.homoGraph([8, 4], start_graph=True, in_dim=30).to(d)
    ei = torch.randint(0, 30, (2, 100), device=d)
    m(None, ei)
    with pytest.raises(RuntimeError, match="Cached 100 number of edges, but found 90"):
        m(None, ei[:, :90])                                      # layers.py:75-81
    m.conv_list[0].cached_result = None                          # PyG idiom: drop the cache
    m(None, ei[:, :90])
    with pytest.raises(RuntimeError):
        gb.myGCN(4, 4)(torch.randn(3, 4), torch.zeros(2, 2, dtype=torch.long))   # CPU tensors: no fallback
    with pytest.raises(AssertionError):
        gb.homoGraph([4, 4], multi_relational=True, n_rela=2).to(d)(torch.randn(5, 4, device=d), ei)
    before = gb.launch_count()
    m(None, ei[:, :90])
    assert gb.launch_count() > before
    me = gb.install_as_gripnet()
    from gripnet.layers import homoGraph as hg          # noqa: E402
    assert hg is gb.homoGraph and me is gb


def test_cuda_graph_capture_of_a_full_step():
    """The whole fwd+loss+bwd step is capturable (no syncs / allocations outside torch's pool)."""
    from gripnet_b200.pipelines import PoseModel, load_flat_params, to_device
    g = synth.pose_small()
    p = synth.pose_params(g)
    m = load_flat_params(PoseModel(g["n_g"], g["n_d"], g["n_rel"]), p).to(_dev())
    data = to_device(g, _dev())
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            m.zero_grad(set_to_none=True)
            loss, *_ = m(data)
            loss.backward()
    torch.cuda.current_stream().wait_stream(s)
    eager = {k: v.grad.clone() for k, v in m.named_parameters() if v.grad is not None}
    eager_loss = loss.detach().clone()
    graph = torch.cuda.CUDAGraph()
    m.zero_grad(set_to_none=True)
    with torch.cuda.graph(graph):
        static_loss, *_ = m(data)
        static_loss.backward()
    for _ in range(2):
        graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(static_loss.detach(), eager_loss)
    for k, v in m.named_parameters():
        if v.grad is not None:
            assert torch.equal(v.grad, eager[k]), k


def test_encoder_rgcn_forward_backward_matches_the_port():
    """``gripnet.encoder.RGCN`` (reference ``encoder.py:6-25``; its ``forward`` reads a misnamed attribute there,
    so the check is against what it evidently computes: project by the embedding, then two myRGCN layers with
    no activation in between) and ``ops.MatMul`` — forward and every gradient against the CPU port."""
    import gripnet_b200 as gb
    from oracle import port, synth
    d = torch.device("cuda:0")
    g = synth.pose_small()
    n, r = g["n_d"], g["n_rel"]
    torch.manual_seed(5)
    enc = gb.encoder.RGCN(24, 16, 12, 8, r, 4).to(d)
    x = torch.randn(n, 24)
    xc = x.to(d).requires_grad_(True)
    ei, et, rl = g["dd_edge_index"].to(d), g["dd_edge_type"].to(d), g["dd_range_list"]
    out = enc(xc, ei, et, rl)
    gvec = torch.linspace(-1, 1, out.numel()).view_as(out)
    (out * gvec.to(d)).sum().backward()
    p = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in enc.named_parameters()}
    xr = x.clone().requires_grad_(True)
    h = xr @ p["embedding"]
    h = port.rgcn_conv(h, p["rgcn1.basis"], p["rgcn1.att"], p["rgcn1.root"], None, g["dd_edge_index"], rl)
    ref = port.rgcn_conv(h, p["rgcn2.basis"], p["rgcn2.att"], p["rgcn2.root"], None, g["dd_edge_index"], rl)
    (ref * gvec).sum().backward()
    assert rel_err(out, ref) < 1e-5
    assert rel_err(xc.grad, xr.grad) < 1e-5
    for k, v in enc.named_parameters():
        assert rel_err(v.grad, p[k].grad) < 1e-5, k
