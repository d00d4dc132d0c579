"""Whole training epochs as ONE CUDA graph.

``PoseTrainer`` is ``train(epoch)`` of ``GripNet-pose.py:113-166``: draw the epoch's negatives
(``utils.negative_sampling``, ``:131``), forward, loss, ``loss.backward()``, ``optimizer.step()`` (Adam,
``:104,146``) and the per-relation AUPRC / AUROC / AP record (``:148-164``) — captured once, replayed per epoch.
Nothing crosses PCIe inside an epoch; the caller reads ``loss`` / ``record`` when it wants them.
``NodeTrainer`` is the node-classification counterpart (``GripNet-aminer.py:120-147``: forward, NC loss,
backward, Adam, micro / macro F1 of the arg-max predictions).
"""
import torch

from . import metrics, streams
from .capture import CapturedStep
from .optim import Adam
from .utils import NegativeSampler


class _EagerStep:
    """The same epoch launched kernel by kernel (profilers cannot follow a graph capture)."""

    def __init__(self, fn, params, post):
        self.fn, self.params, self.post, self.launches_per_replay = fn, params, post, 0

    def replay(self):
        from . import _lib
        for p in self.params:
            p.grad = None
        before = _lib.launch_count()
        self.outputs = self.fn()
        self.outputs[0].backward()
        self.post()
        self.launches_per_replay = _lib.launch_count() - before
        return self.outputs


class _Trainer:
    def _capture(self, fn, model, lr, warmup, post_extra=None, eager=False):
        params = [p for p in model.parameters() if p.requires_grad]
        self.optimizer = Adam(params, lr=lr)
        start = [p.detach().clone() for p in params]

        def post():
            self.optimizer.step()
            if post_extra is not None:
                post_extra()

        if eager:
            self.step = _EagerStep(fn, params, post)
            self.epoch = 0
            return
        self.step = CapturedStep(fn, params, warmup=warmup, post_backward=post)
        # warm-up and capture ran real optimiser steps: rewind parameters, moments and the step counter
        with torch.no_grad():
            for p, s in zip(params, start):
                p.copy_(s)
            for m, v in zip(self.optimizer.exp_avg, self.optimizer.exp_avg_sq):
                m.zero_()
                v.zero_()
            self.optimizer._step.zero_()
        self.epoch = 0
        torch.cuda.synchronize()

    @property
    def launches_per_epoch(self):
        return int(self.step.launches_per_replay)


class PoseTrainer(_Trainer):
    """``model``: ``pipelines.PoseModel``; ``data``: its device dict (``gg_edge_index`` … ``dd_range_list``)."""

    def __init__(self, model, data, lr=0.01, seed=1111, typed_negatives=False, with_metrics=True, warmup=3,
                 eager=False):
        ei = data["dd_edge_index"]
        n_d = int(data["n_d"])
        self.range_list = data["dd_range_list"].to(ei.device).contiguous()
        self.sampler = NegativeSampler(ei, n_d, self.range_list if typed_negatives else None, seed=seed)
        self.neg_edge_index = torch.empty_like(ei)
        n_rel = int(self.range_list.size(0))
        self.record = torch.full((3, n_rel), float("nan"), dtype=torch.float64, device=ei.device) if with_metrics else None
        if with_metrics:                                   # validate + cache the range list outside the capture
            dummy = torch.zeros(ei.size(1), device=ei.device)
            metrics.lp_metrics(dummy, dummy, self.range_list, out=self.record)
        outs = {}

        # the evaluation record needs the forward's scores only: its kernels (rank keys, two radix sorts, the
        # per-relation walk) run on a background branch next to the backward pass and Adam, joined at the epoch's end
        def fn():
            self.sampler.sample(out=self.neg_edge_index)
            outs["o"] = model(data, self.neg_edge_index)
            if with_metrics:
                _, _, pos_score, neg_score = outs["o"]
                outs["br"] = streams.Branch(background=True)
                with outs["br"](pos_score, neg_score):
                    metrics.lp_metrics(pos_score, neg_score, self.range_list, out=self.record)
            return outs["o"]

        def post_extra():
            if with_metrics:
                outs.pop("br").join()

        self._capture(fn, model, lr, warmup, post_extra, eager)
        self.sampler.state.zero_()                         # epoch 0 draws the sampler's first negatives

    def train_epoch(self):
        """One epoch; returns the (device) loss.  ``z`` / ``pos_score`` / ``neg_score`` / ``record`` /
        ``neg_edge_index`` hold this epoch's values until the next call."""
        self.loss, self.z, self.pos_score, self.neg_score = self.step.replay()
        self.epoch += 1
        return self.loss


class NodeTrainer(_Trainer):
    """``model(data) -> (loss, z, score)`` (``pipelines.AminerModel`` / ``FreebaseDModel`` / ``ChainModel``)."""

    def __init__(self, model, data, n_class, lr=0.01, with_metrics=True, warmup=3, eager=False):
        dev = data["train_node_class"].device
        self.f1 = torch.full((3,), float("nan"), dtype=torch.float64, device=dev) if with_metrics else None
        outs = {}

        def fn():
            outs["o"] = model(data)
            if with_metrics:                       # needs the forward's scores only: next to the backward pass
                score = outs["o"][2].detach()
                outs["br"] = streams.Branch(background=True)
                with outs["br"](score):
                    self.pred = metrics.argmax_rows(score)
                    metrics.nc_metrics(data["train_node_class"], self.pred, n_class, out=self.f1)
            return outs["o"]

        def post_extra():
            if with_metrics:
                outs.pop("br").join()

        self._capture(fn, model, lr, warmup, post_extra, eager)

    def train_epoch(self):
        self.loss, self.z, self.score = self.step.replay()
        self.epoch += 1
        return self.loss
