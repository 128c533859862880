"""Per-step drift between the training kernels and the torch fp32 oracle on the golden batches: max |dw| per parameter
group, how many entries moved apart by more than lr (= a sign flip of Adagrad's first step), gradient error per step."""
import sys

import numpy as np
import torch

sys.path[:0] = [".", "tests"]
from oracle.train_oracle import TorchBackend
from pygent_b200.algorithms.mbrl import DynamicsTrainer
from train_cases import dataset, fresh_net


def main():
    torch.set_num_threads(8)
    D = dataset()
    net_o, net_k = fresh_net(), fresh_net()
    to = DynamicsTrainer(net_o, backend=TorchBackend(net_o))
    tk = DynamicsTrainer(net_k, device="cuda")
    for t in (to, tk):
        t.set_moments(D)
        t.set_data(D)
    off, groups = 0, {}
    for name, p in net_o.named_parameters():
        groups[name] = slice(off, off + p.numel())
        off += p.numel()
    step = 0
    for epoch in range(2):
        for idx in D.batch_indices(512, shuffle=False):
            to.backend.grad(idx, len(idx))
            tk.backend.grad(idx, len(idx))
            go, gk = to.backend.bucket.numpy().copy(), tk.backend.bucket.cpu().numpy()
            line = "step %d loss %.8f / %.8f |" % (step, gk[-1], go[-1])
            for name, sl in groups.items():
                line += " %s g %.1e" % (name.replace("layer", "L").replace(".weight", "W").replace(".bias", "b"), np.max(np.abs(gk[sl] - go[sl])) / np.max(np.abs(go[sl])))
            print(line)
            to.backend.update(1e-3, 1e-3)
            tk.backend.update(1e-3, 1e-3)
            po = torch.cat([p.detach().reshape(-1) for p in net_o.parameters()]).numpy()
            pk = tk.backend.k.params.cpu().numpy()
            line = "        weights |"
            for name, sl in groups.items():
                d = np.abs(pk[sl] - po[sl])
                line += " %s max %.1e n(>1e-4) %d" % (name.replace("layer", "L").replace(".weight", "W").replace(".bias", "b"), d.max(), int((d > 1e-4).sum()))
            print(line)
            step += 1


if __name__ == "__main__":
    main()
