"""CPU: pin the oracle (port + float64 dense) to the reference's golden fixtures.

The fixtures are outputs of the unmodified reference (tests/golden/make_golden.py).
Tolerance: fp32 results within 1e-5 relative (max|a-b|/max|b| per tensor), the
tolerance BASELINE.json's north_star states; integer/structure results bit-exact.
"""
import numpy as np
import pytest
import torch

from golden_util import load_case, load_grouped, rel_err
from oracle import dense64, port

TOL = 1e-5


def _leafify(p):
    return {k: v.clone().requires_grad_(True) for k, v in p.items()}


def _check_grads(p, golden_grads, tol=TOL):
    for k, g in golden_grads.items():
        assert p[k].grad is not None, k
        assert rel_err(p[k].grad, g) < tol, (k, rel_err(p[k].grad, g))


# ------------------------------------------------------------------ a1
@pytest.mark.parametrize("name", ["kat6", "loops_w", "loops_improved", "isolated", "empty"])
def test_gcn_norm_matches_reference(name):
    c = load_grouped("gcn_norm")[name]
    ei, nrm = port.gcn_norm(c["edge_index"], int(c["num_nodes"]), c.get("edge_weight"), bool(c["improved"]))
    assert torch.equal(ei, c["out_edge_index"])                      # structure: bit-exact
    assert torch.equal(nrm, c["out_norm"])                           # same op order on CPU: bit-exact


def test_kat6_literal():
    """The hand-checked example of SURVEY.md §8 a1."""
    ei, nrm = port.gcn_norm(torch.tensor([[0, 1, 1, 2, 2, 0], [1, 1, 2, 0, 0, 3]]), 4)
    assert ei.tolist() == [[0, 1, 2, 2, 0, 0, 1, 2, 3], [1, 2, 0, 0, 3, 0, 1, 2, 3]]
    o = port.gcn_csr_oracle(np.array([[0, 1, 1, 2, 2, 0], [1, 1, 2, 0, 0, 3]]), 4)
    assert o["indeg"].tolist() == [3, 2, 2, 2]
    assert o["rowptr"].tolist() == [0, 3, 5, 7, 9]
    assert o["col"].tolist() == [2, 2, 0, 0, 1, 1, 2, 0, 3]       # in-edges in original order, loop last
    assert o["perm"].tolist() == [2, 3, 5, 0, 6, 1, 7, 4, 8]


def test_csr_oracle_properties():
    rs = np.random.RandomState(0)
    ei = rs.randint(0, 30, (2, 300))
    o = port.gcn_csr_oracle(ei, 30, rs.uniform(0.5, 1.5, 300).astype(np.float32))
    src, dst = o["edge_index_aug"]
    for rp, perm, key in ((o["rowptr"], o["perm"], dst), (o["rowptr_t"], o["perm_t"], src)):
        assert sorted(perm.tolist()) == list(range(len(perm)))
        assert (np.diff(key[perm]) >= 0).all()
        for i in range(30):
            seg = perm[rp[i]:rp[i + 1]]
            assert (key[seg] == i).all() and (np.diff(seg) > 0).all()   # stable
    # loop is the last entry of each dst row
    assert (o["col"][o["rowptr"][1:] - 1] == np.arange(30)).all()


# ------------------------------------------------------------------ full models
@pytest.mark.parametrize("name", ["pose_small", "pose_small_weighted"])
def test_pose_port_matches_reference(name):
    c = load_case(name)
    p = _leafify(c["p"])
    loss, z, pos, neg = port.pose_forward(p, c["in"])
    loss.backward()
    assert rel_err(z, c["out"]["z"]) < TOL
    assert rel_err(pos, c["out"]["pos"]) < TOL and rel_err(neg, c["out"]["neg"]) < TOL
    assert rel_err(loss, c["out"]["loss"]) < TOL
    _check_grads(p, c["grad"])


def test_aminer_port_matches_reference():
    c = load_case("aminer_small")
    p = _leafify(c["p"])
    loss, z, score = port.aminer_forward(p, c["in"])
    loss.backward()
    assert rel_err(z, c["out"]["z"]) < TOL and rel_err(score, c["out"]["score"]) < TOL
    assert rel_err(loss, c["out"]["loss"]) < TOL
    _check_grads(p, c["grad"])


def test_freebase_d_port_matches_reference():
    c = load_case("freebase_d_small")
    p = _leafify(c["p"])
    loss, z, score = port.freebase_d_forward(p, c["in"])
    loss.backward()
    assert rel_err(z, c["out"]["z"]) < TOL and rel_err(score, c["out"]["score"]) < TOL
    assert rel_err(loss, c["out"]["loss"]) < TOL
    _check_grads(p, c["grad"])


# ------------------------------------------------------------------ module variants
def _split(c):
    p = {k[2:]: v for k, v in c.items() if k.startswith("p.")}
    g = {k[5:]: v for k, v in c.items() if k.startswith("grad.")}
    return p, g


def _weights_like(out):
    return torch.linspace(-1, 1, out.numel()).view_as(out)


@pytest.mark.parametrize("tag,mod,relu", [("add_eq", "add", True), ("add_down", "add", True),
                                          ("cat_norelu", "cat", False)])
def test_inter_variants(tag, mod, relu):
    c = load_grouped("variants")["inter_" + tag]
    p0, g = _split(c)
    p = _leafify(p0)
    x = c["x"].clone().requires_grad_(True)
    n_t = p0["target_feat"].shape[0]
    out = port.inter_forward(p, x, c["edge_index"], n_t, c.get("edge_weight"), if_relu=relu, mod=mod)
    (out * _weights_like(out)).sum().backward()
    assert rel_err(out, c["out"]) < TOL and rel_err(x.grad, c["grad_x"]) < TOL
    _check_grads(p, g)
    d = dense64.inter(p0, c["x"], c["edge_index"], n_t, c.get("edge_weight"), if_relu=relu, mod=mod)
    assert rel_err(d, c["out"]) < TOL


def test_homo_nocat_and_dense():
    c = load_grouped("variants")["homo_nocat"]
    p0, g = _split(c)
    p = _leafify(p0)
    x = c["x"].clone().requires_grad_(True)
    out = port.homo_forward(p, x, c["edge_index"], c["edge_weight"])
    (out * _weights_like(out)).sum().backward()
    assert rel_err(out, c["out"]) < TOL and rel_err(x.grad, c["grad_x"]) < TOL
    _check_grads(p, g)
    assert rel_err(dense64.homo(p0, c["x"], c["edge_index"], c["edge_weight"]), c["out"]) < TOL


def test_rgcn_variants():
    v = load_grouped("variants")
    c = v["rgcn2"]
    p0, g = _split(c)
    p = _leafify(p0)
    x = c["x"].clone().requires_grad_(True)
    out = port.homo_forward(p, x, c["edge_index"], edge_type=c["edge_type"], range_list=c["range_list"],
                            if_catout=True, multi_relational=True)
    (out * _weights_like(out)).sum().backward()
    assert rel_err(out, c["out"]) < TOL and rel_err(x.grad, c["grad_x"]) < TOL
    _check_grads(p, g)
    d = dense64.homo(p0, c["x"], c["edge_index"], range_list=c["range_list"], if_catout=True, multi_relational=True)
    assert rel_err(d, c["out"]) < TOL

    c = v["rgcn_bias"]
    p0, g = _split(c)
    p = _leafify(p0)
    x = c["x"].clone().requires_grad_(True)
    out = port.rgcn_conv(x, p["basis"], p["att"], p["root"], p["bias"], c["edge_index"], c["range_list"])
    (out * _weights_like(out)).sum().backward()
    assert rel_err(out, c["out"]) < TOL and rel_err(x.grad, c["grad_x"]) < TOL
    _check_grads(p, g)


def test_gcn_improved():
    c = load_grouped("variants")["gcn_improved"]
    p0, g = _split(c)
    p = _leafify(p0)
    x = c["x"].clone().requires_grad_(True)
    ei, nrm = port.gcn_norm(c["edge_index"], x.shape[0], None, improved=True)
    out = port.gcn_conv(x, p["weight"], None, ei, nrm)
    (out * _weights_like(out)).sum().backward()
    assert rel_err(out, c["out"]) < TOL and rel_err(x.grad, c["grad_x"]) < TOL
    _check_grads(p, g)
    a = dense64.gcn_adj(c["edge_index"], x.shape[0], improved=True)
    assert rel_err(dense64.gcn(c["x"], p0["weight"], None, a), c["out"]) < TOL


def test_decoder_variants():
    v = load_grouped("variants")
    c = v["dmt_raw"]
    z = c["z"].clone().requires_grad_(True)
    w = c["p.weight"].clone().requires_grad_(True)
    out = port.distmult(z, w, c["edge_index"], c["edge_type"], sigmoid=False)
    (out * _weights_like(out)).sum().backward()
    assert rel_err(out, c["out"]) < TOL and rel_err(z.grad, c["grad_z"]) < TOL
    assert rel_err(w.grad, c["grad.weight"]) < TOL
    assert rel_err(dense64.distmult(c["z"], c["p.weight"], c["edge_index"], c["edge_type"], False), c["out"]) < TOL
    for tag, sm in (("mcip_raw", False), ("mcip_softmax", True)):
        c = v[tag]
        z = c["z"].clone().requires_grad_(True)
        w = c["p.weight"].clone().requires_grad_(True)
        out = port.multiclass(z, w, c["node_list"], softmax=sm)
        (out * _weights_like(out)).sum().backward()
        assert rel_err(out, c["out"]) < TOL and rel_err(z.grad, c["grad_z"]) < TOL
        assert rel_err(w.grad, c["grad.weight"]) < TOL
        assert rel_err(dense64.multiclass(c["z"], c["p.weight"], c["node_list"], sm), c["out"]) < TOL


def test_dense64_agrees_on_pose_stages():
    """Independent formulation vs the reference outputs, stage by stage (fp32 tolerance)."""
    c = load_case("pose_small_weighted")
    p, g = c["p"], c["in"]
    sub = lambda pre: {k[len(pre):]: v for k, v in p.items() if k.startswith(pre)}
    z_gg = dense64.homo(sub("gg."), None, g["gg_edge_index"], g["gg_edge_weight"], if_catout=True)
    assert rel_err(z_gg, c["out"]["z_gg"]) < TOL
    z_gd = dense64.inter(sub("gd."), z_gg, g["gd_edge_index"], g["n_d"])
    assert rel_err(z_gd, c["out"]["z_gd"]) < TOL
    z = dense64.homo(sub("dd."), z_gd, g["dd_edge_index"], range_list=g["dd_range_list"], if_catout=True,
                     multi_relational=True)
    assert rel_err(z, c["out"]["z"]) < TOL
    pos = dense64.distmult(z, p["dmt.weight"], g["dd_edge_index"], g["dd_edge_type"])
    assert rel_err(pos, c["out"]["pos"]) < TOL


def test_pinned_relu_pattern_is_the_identity_on_the_oracles_own_pattern():
    """oracle/port.py:_relu — replaying the branch pattern the oracle itself took changes nothing
    (this is the hook the full-size GPU gradient tests use)."""
    from oracle import port, synth
    g = synth.pose_small()
    p = synth.pose_params(g)
    sub = port._sub
    with torch.no_grad():
        z_gg = port.homo_forward(sub(p, "gg."), None, g["gg_edge_index"], if_catout=True)
        z_gd = port.inter_forward(sub(p, "gd."), z_gg, g["gd_edge_index"], g["n_d"], mod="cat", if_relu=True)
        z_dd = port.homo_forward(sub(p, "dd."), z_gd, g["dd_edge_index"], edge_type=g["dd_edge_type"],
                                 range_list=g["dd_range_list"], if_catout=True, multi_relational=True)
    res = []
    for patterns in (None, {"gg": z_gg, "gd": z_gd, "dd": z_dd}):
        pl = {k: v.clone().requires_grad_(True) for k, v in p.items()}
        out = port.pose_forward(pl, g, patterns=patterns)
        out[0].backward()
        res.append((out, pl))
    assert torch.equal(res[0][0][1], z_dd)
    assert torch.allclose(res[0][0][0], res[1][0][0], rtol=0, atol=0)
    for k in p:
        a, b = res[0][1][k].grad, res[1][1][k].grad
        assert (a is None) == (b is None)
        if a is not None:
            assert torch.equal(a, b), k
