"""Data-parallel dynamics training on real GPUs, one rank per GPU (torchrun): the NCCL all-reduce path and the fused
NVLink peer-memory path (pg_mlp_train_update_p2p) against each other and against the single-GPU full-batch gradient.
Rank 0 prints one JSON line.   torchrun --nproc-per-node N tools/dist_train_check.py [steps]"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__)))]
sys.path.insert(1, os.path.join(sys.path[0], "tests"))
from pygent_b200.algorithms.mbrl import DynamicsTrainer  # noqa: E402
from train_cases import dataset, fresh_net  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    D = dataset()
    out = {"world": world}
    trainers = {}
    for ex in ("nccl", "p2p"):
        net = fresh_net()
        try:
            tr = DynamicsTrainer(net, dyn_lr=1e-3, weight_decay=1e-3, device=dev, exchange=ex)
        except Exception as e:  # noqa: BLE001
            out[ex + "_error"] = "%s: %s" % (type(e).__name__, e)
            continue
        tr.set_moments(D)
        tr.set_data(D)
        trainers[ex] = tr
    # (1) the all-reduced bucket of the first step == the full-batch gradient computed on one GPU
    idx = np.arange(512)
    ref = DynamicsTrainer(fresh_net(), dyn_lr=1e-3, weight_decay=1e-3, device=dev, exchange="nccl")
    ref.set_moments(D)
    ref.set_data(D)
    ref.backend.grad(idx, 512)  # the whole batch on this rank
    full = ref.backend.bucket.clone()
    tn = trainers["nccl"]
    tn.backend.grad(idx[rank::world], 512)
    dist.all_reduce(tn.backend.bucket)
    err = float((tn.backend.bucket - full).abs().max() / full.abs().max())
    out["allreduced_vs_full_batch_rel"] = err
    tn.backend.update(1e-3, 1e-3)
    if "p2p" in trainers:
        tp = trainers["p2p"]
        lp = tp.train_step(idx)
        out["p2p_loss_vs_full"] = float(abs(float(lp) - float(full[-1])) / abs(float(full[-1])))
        d = (tp.backend.k.params - tn.backend.k.params).abs()
        out["p2p_vs_nccl_weights_after_1_step"] = {"max": float(d.max()), "n_gt_1e-6": int((d > 1e-6).sum())}
    # (2) several steps; ranks must stay bit-identical
    batches = D.batch_indices(512, shuffle=False)
    for name, tr in trainers.items():
        losses = []
        for ep in range(2):
            for b in batches:
                losses.append(float(tr.train_step(b)))
        chk = tr.backend.k.params.double().sum().reshape(1)
        every = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(every, chk)
        out[name + "_losses"] = losses
        out[name + "_ranks_identical"] = bool(all(float(e) == float(every[0]) for e in every))
    if "p2p" in trainers:
        out["p2p_status"] = trainers["p2p"].backend.k.p2p_status()
    # (3) timing, device events, max over ranks
    idx_all = [np.arange(i * 512, (i + 1) * 512) % 1200 for i in range(8)]
    for name, tr in trainers.items():
        for i in range(10):
            tr.train_step(idx_all[i % 8])
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(steps):
            tr.train_step(idx_all[i % 8])
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[name + "_us_per_step"] = float(t[0]) * 1e3 / steps
    if rank == 0:
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
