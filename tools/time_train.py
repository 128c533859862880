"""Timing of the dynamics training step on one GPU: kernels (pg_mlp_train_step) vs the same step in eager torch (autograd +
torch.optim.Adagrad, cuBLAS) -- the round-1 implementation.   python tools/time_train.py [rows]"""
import sys
import time

import numpy as np
import torch

sys.path[:0] = [".", "tests"]
from oracle.train_oracle import TorchBackend
from pygent_b200.algorithms.mbrl import DynamicsTrainer
from train_cases import dataset, fresh_net


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    D = dataset()
    net = fresh_net()
    tk = DynamicsTrainer(net, device="cuda", batch_size=rows)
    tk.set_moments(D)
    tk.set_data(D)
    k = tk.backend.k
    idx = torch.as_tensor(np.arange(rows, dtype=np.int32) % 1200, device="cuda")
    for _ in range(20):
        k.step(idx, 1e-3, 1e-3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 500
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        k.step(idx, 1e-3, 1e-3)
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    print("kernels: %.1f us/step device, %.1f us/step wall, loss %.6f" % (e0.elapsed_time(e1) * 1e3 / n, wall * 1e6 / n, float(k.bucket[-1])))
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10):
            k.step(idx, 1e-3, 1e-3)
    g.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(50):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print("kernels, CUDA graph of 10 steps: %.1f us/step" % (e0.elapsed_time(e1) * 1e3 / 500))
    # eager torch on the GPU (what round 1 shipped)
    net2 = fresh_net().cuda()
    opt = torch.optim.Adagrad(net2.parameters(), lr=1e-3, weight_decay=1e-3)
    o = torch.randn(rows, 5, device="cuda"); u = torch.randn(rows, 1, device="cuda"); y = torch.randn(rows, 2, device="cuda")
    for _ in range(20):
        opt.zero_grad(); loss = ((net2(o, u) - y) ** 2).mean(); loss.backward(); opt.step()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(200):
        opt.zero_grad(); loss = ((net2(o, u) - y) ** 2).mean(); loss.backward(); opt.step()
    e1.record()
    torch.cuda.synchronize()
    print("eager torch (cuBLAS + autograd + optim.Adagrad): %.1f us/step" % (e0.elapsed_time(e1) * 1e3 / 200))


if __name__ == "__main__":
    main()
