"""GPU: whole training epochs replayed from one CUDA graph (training.PoseTrainer / NodeTrainer) against an
eager loop written like the reference script (GripNet-pose.py:113-166): same modules, torch.optim.Adam, the
negatives of the same counter-based sampler, sklearn metrics on the host."""
import numpy as np
import pytest
import torch

from oracle import metrics as om

pytestmark = pytest.mark.gpu


def _pose(seed=1111):
    from gripnet_b200.pipelines import PoseModel, to_device
    from gripnet_b200.synthetic import pose_small
    g = pose_small(seed)
    dev = torch.device("cuda:0")
    torch.manual_seed(7)
    model = PoseModel(g["n_g"], g["n_d"], g["n_rel"]).to(dev)
    return g, model, to_device(g, dev)


def test_pose_trainer_matches_an_eager_reference_style_loop():
    import copy
    from gripnet_b200.training import PoseTrainer
    from gripnet_b200.utils import NegativeSampler
    g, model, data = _pose()
    twin = copy.deepcopy(model)
    epochs = 6
    trainer = PoseTrainer(model, data, lr=0.01, seed=5)
    assert trainer.optimizer.step_count == 0 and trainer.sampler.epoch == 0
    for a, b in zip(model.parameters(), twin.parameters()):
        assert torch.equal(a, b), "capture must leave the parameters untouched"

    # the reference loop: zero_grad, forward, negative_sampling, scores, loss, backward, optimizer.step, metrics
    opt = torch.optim.Adam(twin.parameters(), lr=0.01, foreach=False, fused=False)
    sampler = NegativeSampler(data["dd_edge_index"], g["n_d"], None, seed=5)
    rl = g["dd_range_list"].numpy()
    for ep in range(epochs):
        loss = trainer.train_epoch()
        opt.zero_grad()
        neg = sampler.sample()
        ref_loss, _, pos_score, neg_score = twin(data, neg)
        ref_loss.backward()
        opt.step()
        assert torch.equal(trainer.neg_edge_index, neg), "same counter-based draw in and out of the graph"
        assert abs(float(loss) - float(ref_loss)) < 2e-5 * abs(float(ref_loss)), (ep, float(loss), float(ref_loss))
        want = om.lp_record(pos_score.detach().cpu().numpy(), neg_score.detach().cpu().numpy(), rl)
        got = trainer.record.cpu().numpy()
        # the metrics are exact functions of the scores; the scores differ by fp32 round-off between the two
        # optimisers, which can swap neighbouring ranks: compare on the trainer's own scores exactly, and the
        # two runs loosely
        own = om.lp_record(trainer.pos_score.detach().cpu().numpy(), trainer.neg_score.detach().cpu().numpy(), rl)
        np.testing.assert_allclose(got, own, rtol=1e-12)
        np.testing.assert_allclose(got, want, atol=2e-3)
    assert trainer.epoch == epochs and trainer.optimizer.step_count == epochs and trainer.sampler.epoch == epochs
    for (k, a), b in zip(model.named_parameters(), twin.parameters()):
        err = float((a - b).abs().max() / b.abs().max())
        assert err < 5e-3, (k, err)     # Adam normalises every gradient component to ~lr: one whose value is a
        #                                 cancellation residue can take a different sign in the two runs (the
        #                                 per-epoch losses above are the strict check)
    assert trainer.launches_per_epoch > 40
    losses = []
    for _ in range(20):
        losses.append(float(trainer.train_epoch()))
    assert losses[-1] < losses[0], "training reduces the loss"


def test_pose_trainer_is_deterministic():
    from gripnet_b200.training import PoseTrainer
    runs = []
    for _ in range(2):
        _, model, data = _pose()
        t = PoseTrainer(model, data, lr=0.01, seed=9, with_metrics=False)
        for _ in range(4):
            loss = t.train_epoch()
        runs.append((float(loss), [p.detach().clone() for p in model.parameters()]))
    assert runs[0][0] == runs[1][0]
    for a, b in zip(runs[0][1], runs[1][1]):
        assert torch.equal(a, b)


def test_node_trainer_matches_an_eager_loop():
    import copy
    from sklearn import metrics as skm
    from gripnet_b200.pipelines import AminerModel, to_device
    from gripnet_b200.synthetic import aminer_small
    from gripnet_b200.training import NodeTrainer
    g = aminer_small()
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    model = AminerModel(g["n_p"], g["n_a"], g["n_class"], pp=(32, 16, 16), pa=(16, 16), aa_hid=(32, 8)).to(dev)
    twin = copy.deepcopy(model)
    data = to_device(g, dev)
    trainer = NodeTrainer(model, data, g["n_class"], lr=0.01)
    opt = torch.optim.Adam(twin.parameters(), lr=0.01, foreach=False, fused=False)
    y = g["train_node_class"].numpy()
    for ep in range(5):
        loss = trainer.train_epoch()
        opt.zero_grad()
        ref_loss, _, score = twin(data)
        ref_loss.backward()
        opt.step()
        assert abs(float(loss) - float(ref_loss)) < 2e-5 * abs(float(ref_loss)), (ep, float(loss), float(ref_loss))
        pred = trainer.score.detach().argmax(1)
        assert torch.equal(trainer.pred, pred)
        micro, macro, acc = trainer.f1.cpu().numpy()
        p = pred.cpu().numpy()
        assert micro == pytest.approx(skm.f1_score(y, p, average="micro"), rel=1e-12)
        assert macro == pytest.approx(skm.f1_score(y, p, average="macro"), rel=1e-12)
        assert acc == pytest.approx(skm.accuracy_score(y, p), rel=1e-12)
