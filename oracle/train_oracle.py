"""Torch fp32 restatement of the reference's dynamics-training step (TEST INFRASTRUCTURE, NOT PRODUCT).

Follows pygent/algorithms/mbrl.py line by line with torch autograd on the CPU:
  training_data   :471-485   o_, u, dx = x - x_ (velocity half with sparse_dyn) standardised with the data-set moments
  training        :450-469   nn.MSELoss, zero_grad, backward, optimiser step, batch by batch
  optimiser       :71-73     torch.optim.Adagrad(lr = dyn_lr, weight_decay = weight_decay)
Pinned by tests/golden/mlp_training.npz (oracle/make_golden_train.py runs the reference's own MBRL.set_moments /
MBRL.training_data / Adagrad on the same batches): tests/test_training.py::test_training_oracle_matches_reference.

`TorchBackend` speaks the backend protocol of pygent_b200.algorithms.mbrl.DynamicsTrainer (set_moments / set_data / grad /
update / bucket), so the CPU tests can drive the product's host logic -- batch slicing, ONE bucket all-reduce per step on a
gloo group -- without a GPU; the product itself only ever constructs the CUDA backend (pygent_b200.mlp.TrainKernels).
"""
import numpy as np
import torch


class TorchBackend:
    def __init__(self, net, sparse_dyn=True):
        self.net = net
        self.sparse_dyn = sparse_dyn
        self.plist = list(net.parameters())
        self.count = sum(p.numel() for p in self.plist)
        self.bucket = torch.zeros(self.count + 1, dtype=torch.float32)
        off = 0
        for p in self.plist:  # every parameter's .grad is a view into the flat bucket
            p.grad = self.bucket[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.optim = None
        self.mom = None
        self.data = None
        self.launches = 0

    def set_moments(self, o_mean, o_var, u_mean, u_var, dx_mean, dx_var):
        t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float32)).reshape(1, -1)
        self.mom = dict(o_mean=t(o_mean), o_var=t(o_var), u_mean=t(u_mean), u_var=t(u_var), dx_mean=t(dx_mean), dx_var=t(dx_var))

    def set_data(self, col):
        self.data = {k: torch.as_tensor(np.asarray(col[k], dtype=np.float32)) for k in ("x_", "x", "o_", "u")}

    def outputs(self, idx):
        """(dx_norm, fOutputs_norm) of the rows idx, as MBRL.training_data returns them."""
        d, m = self.data, self.mom
        idx = torch.as_tensor(np.asarray(idx, dtype=np.int64))
        x_, x, o_, u = d["x_"][idx], d["x"][idx], d["o_"][idx], d["u"][idx]
        o_n = (o_ - m["o_mean"]) / m["o_var"]
        u_n = (u - m["u_mean"]) / m["u_var"]
        dx = x - x_
        if self.sparse_dyn:
            dx = dx[:, dx.shape[1] // 2:]
        return (dx - m["dx_mean"]) / m["dx_var"], self.net(o_n, u_n)

    def grad(self, idx, n_total):
        self.bucket.zero_()
        if len(idx):
            dx_n, f = self.outputs(idx)
            # nn.MSELoss over the whole batch = sum of squared errors / (rows * outputs); this slice's share of the sum
            loss = ((f - dx_n) ** 2).sum() / float(n_total * dx_n.shape[1])
            loss.backward()
            self.bucket[-1] = loss.detach()

    def update(self, lr, weight_decay, eps=1e-10):
        if self.optim is None:
            self.optim = torch.optim.Adagrad(self.plist, lr=lr, weight_decay=weight_decay, eps=eps)
        self.optim.step()

    def sync_module(self):
        pass
