#!/usr/bin/env python
"""Node-classification experiment of the reference (``GripNet-aminer.py`` / ``GripNet-freebase-d.py``) on
gripnet_b200.

    python examples/gripnet_nc.py EPOCHS --model aminer datasets/aminer.pt --train label.train --test label.test
    python examples/gripnet_nc.py EPOCHS --model aminer --synthetic
    python examples/gripnet_nc.py EPOCHS --model freebase-d --synthetic

Model wiring: ``GripNet-aminer.py:96-108`` / ``GripNet-freebase-d.py:103-137``; Adam lr 0.01; loss ``:133``;
micro / macro F1 of the arg-max predictions (``:131,137,154-156``).  An epoch is one CUDA-graph replay
(``training.NodeTrainer``).
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def _label_file(path, n_class, utils):
    import numpy as np
    arr = np.loadtxt(path, dtype=np.int64, delimiter="\t").T          # rows: node id, label
    return utils.process_data_multiclass(torch.from_numpy(arr), n_class)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("epochs", type=int)
    ap.add_argument("dataset", nargs="?")
    ap.add_argument("--model", choices=["aminer", "freebase-d"], default="aminer")
    ap.add_argument("--synthetic", action="store_true")
    ap.add_argument("--train")
    ap.add_argument("--test")
    ap.add_argument("--lr", type=float, default=0.01)
    ap.add_argument("--out", default="out/nc")
    args = ap.parse_args()

    from gripnet_b200 import data as gd, metrics, utils
    from gripnet_b200.pipelines import AminerModel, FreebaseDModel, to_device
    from gripnet_b200.training import NodeTrainer
    dev = torch.device("cuda:0")
    torch.manual_seed(1111)
    if args.synthetic or args.dataset is None:
        from gripnet_b200 import synthetic
        g = synthetic.aminer_full() if args.model == "aminer" else synthetic.freebase_d_full()
        test_idx, test_cls = g["train_node_idx"], g["train_node_class"]
    else:
        ds = gd.load(args.dataset)
        ds.train_node_idx, ds.train_node_class, _ = _label_file(args.train, int(ds.n_a_type), utils)
        ds.test_node_idx, ds.test_node_class, _ = _label_file(args.test, int(ds.n_a_type), utils)
        g = gd.nc_inputs(ds, "train")
        test_idx, test_cls = ds.test_node_idx, ds.test_node_class
    g = to_device(g, dev)
    test_idx, test_cls = test_idx.to(dev), test_cls.to(dev)
    if args.model == "aminer":
        model = AminerModel(g["n_p"], g["n_a"], g["n_class"]).to(dev)
    else:
        model = FreebaseDModel(g["n_p"], g["n_q"], g["n_a"], g["n_class"]).to(dev)
    print(model)
    trainer = NodeTrainer(model, g, g["n_class"], lr=args.lr)
    hist = torch.empty(args.epochs, 5, dtype=torch.float64)
    for epoch in range(args.epochs):
        t0 = time.time()
        loss = trainer.train_epoch()
        with torch.no_grad():
            pred = metrics.argmax_rows(model.mcip(trainer.z.detach(), test_idx))
            te = metrics.nc_metrics(test_cls, pred, g["n_class"])
        tr = trainer.f1.cpu()
        te = te.cpu()
        hist[epoch] = torch.tensor([float(loss.detach()), float(tr[0]), float(tr[1]), float(te[0]), float(te[1])])
        print("{:3d}   loss:{:0.4f}   train micro:{:0.4f} macro:{:0.4f}   test micro:{:0.4f} macro:{:0.4f}   "
              "time:{:0.3f}".format(epoch, *hist[epoch].tolist(), time.time() - t0))
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    torch.save({k: v.cpu() for k, v in model.state_dict().items()}, args.out + "-model.pt")
    torch.save({"history": hist}, args.out + "-record.pt")


if __name__ == "__main__":
    main()
