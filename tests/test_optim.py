"""Fused Adam (gn_adam_step): oracle pinned to torch.optim.Adam on CPU; device kernel vs the oracle."""
import numpy as np
import pytest
import torch

from oracle.adam import AdamOracle

TOL = 1e-5      # SURVEY.md §7 tolerance definition: max|a-b| / max|b| per tensor, fp32 bar of the path


def _close(a, b, tol=TOL):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    err = np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30)
    assert err < tol, f"max|a-b|/max|b| = {err:.3e}"


SHAPES = [(37, 16), (5,), (1,), (130, 7), (4096,), (4097,), (3, 3, 5)]


def _problem(seed, steps):
    rs = np.random.RandomState(seed)
    params = [rs.randn(*s).astype(np.float32) for s in SHAPES]
    grads = [[(rs.randn(*s) * 10.0 ** rs.uniform(-4, 1)).astype(np.float32) for s in SHAPES] for _ in range(steps)]
    return params, grads


@pytest.mark.parametrize("kw", [dict(lr=1e-2), dict(lr=1e-3, betas=(0.8, 0.99), eps=1e-6), dict(lr=5e-3, weight_decay=0.1)])
def test_oracle_matches_torch_adam(kw):
    """Pin: the numpy restatement reproduces torch.optim.Adam (single-tensor path) to float32 round-off
    (torch fuses multiply-adds, numpy does not; ulp differences are amplified by m / sqrt(v): 1e-5 relative, the
    fp32 bar of the path)."""
    steps = 12
    params, grads = _problem(0, steps)
    tp = [torch.tensor(p, requires_grad=True) for p in params]
    opt = torch.optim.Adam(tp, foreach=False, **kw)
    orc = AdamOracle(params, **kw)
    for g in grads:
        for p, gi in zip(tp, g):
            p.grad = torch.tensor(gi)
        opt.step()
        orc.step(g)
    for a, b in zip(orc.p, tp):
        _close(a, b.detach().numpy())
    st = opt.state[tp[0]]
    _close(orc.m[0], st["exp_avg"].numpy())
    _close(orc.v[0], st["exp_avg_sq"].numpy())


def test_oracle_skips_parameters_without_gradient():
    params, grads = _problem(1, 1)
    orc = AdamOracle(params, lr=1e-2)
    g = list(grads[0])
    g[2] = None
    orc.step(g)
    assert np.array_equal(orc.p[2], params[2]) and not np.array_equal(orc.p[0], params[0])


# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(lr=1e-2), dict(lr=1e-3, betas=(0.8, 0.99), eps=1e-6), dict(lr=5e-3, weight_decay=0.1)])
def test_device_adam_matches_oracle(kw):
    from gripnet_b200.optim import Adam
    steps = 12
    params, grads = _problem(2, steps)
    dp = [torch.tensor(p, device="cuda", requires_grad=True) for p in params]
    opt = Adam(dp, **kw)
    orc = AdamOracle(params, **kw)
    for g in grads:
        for p, gi in zip(dp, g):
            p.grad = torch.tensor(gi, device="cuda")
        opt.step()
        orc.step(g)
    assert opt.step_count == steps
    for a, b, m, mo, v, vo in zip(orc.p, dp, opt.exp_avg, orc.m, opt.exp_avg_sq, orc.v):
        _close(b.detach().cpu().numpy(), a)     # fp32 round-off (FMA contraction)
        _close(m.cpu().numpy(), mo)
        _close(v.cpu().numpy(), vo)


@pytest.mark.gpu
def test_device_adam_matches_torch_adam_on_device_and_state_dict_round_trip():
    from gripnet_b200.optim import Adam
    params, grads = _problem(3, 6)
    a = [torch.tensor(p, device="cuda", requires_grad=True) for p in params]
    b = [torch.tensor(p, device="cuda", requires_grad=True) for p in params]
    ours, ref = Adam(a, lr=1e-2), torch.optim.Adam(b, lr=1e-2, foreach=False, fused=False)
    for k, g in enumerate(grads):
        for p, q, gi in zip(a, b, g):
            p.grad = torch.tensor(gi, device="cuda")
            q.grad = torch.tensor(gi, device="cuda")
        if k == 3:                      # a parameter without gradient is skipped by both
            a[1].grad = None
            b[1].grad = None
        ours.step()
        ref.step()
        if k == 2:                      # checkpoint -> fresh optimiser -> continue
            sd = ours.state_dict()
            ours = Adam(a, lr=1.0)
            ours.load_state_dict(sd)
    for p, q in zip(a, b):
        # parameter 1 missed one update in torch (its own step counter) — compare the others
        if p is a[1]:
            continue
        _close(p.detach().cpu().numpy(), q.detach().cpu().numpy())
    sd = ours.state_dict()
    assert set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"} and float(sd["state"][0]["step"]) == 6.0


@pytest.mark.gpu
def test_adam_in_a_cuda_graph_advances_every_replay():
    from gripnet_b200.optim import Adam
    params, _ = _problem(4, 1)
    dp = [torch.tensor(p, device="cuda", requires_grad=True) for p in params]
    g_static = [torch.zeros_like(p) for p in dp]
    for p, g in zip(dp, g_static):
        p.grad = g
    opt = Adam(dp, lr=1e-2)
    orc = AdamOracle(params, lr=1e-2)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        with torch.cuda.graph(graph, stream=s):
            opt.step()
    torch.cuda.current_stream().wait_stream(s)
    rs = np.random.RandomState(5)
    for _ in range(5):
        g = [rs.randn(*sh).astype(np.float32) for sh in SHAPES]
        for gs, gi in zip(g_static, g):
            gs.copy_(torch.tensor(gi))
        graph.replay()
        orc.step(g)
    torch.cuda.synchronize()
    assert opt.step_count == 5
    for a, b in zip(orc.p, dp):
        _close(b.detach().cpu().numpy(), a)


@pytest.mark.gpu
def test_adam_rejects_cpu_parameters():
    from gripnet_b200.optim import Adam
    with pytest.raises(RuntimeError):
        Adam([torch.zeros(3, requires_grad=True)])
