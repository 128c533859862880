"""Training of the learned dynamics model: the supervised half of the reference's MBRL
(pygent/algorithms/mbrl.py: `set_moments` :487-509, `training_data` :471-485, `training` :450-469, `pred_loss`
:187-206; optimiser `torch.optim.Adagrad(lr=dyn_lr, weight_decay=weight_decay)` :71-73, MSE on standardised velocity
increments, batch 512).

Data parallel: every rank holds the same data set and the same network; a batch is cut into `world_size` strided
slices, each rank back-propagates its slice of the SUM of squared errors divided by the full batch size, the gradients
live in ONE flat fp32 bucket (255,002 elements) that is all-reduced once per step on the compute stream (NCCL; gloo in
the CPU tests), and every rank applies the identical Adagrad update.  This is the only collective of the learned-
dynamics path besides the final cost/argmin gather.  The dense algebra here is plain library GEMMs
(cuBLAS through torch); the planner's hot path through the network is `pygent_b200.mlp`.
"""
import numpy as np
import torch
import torch.distributed as dist


class DynamicsTrainer(object):
    def __init__(self, nn_dynamics, dyn_lr=1e-3, weight_decay=1e-3, batch_size=512, training_epochs=60, sparse_dyn=True,
                 device=None, process_group=None):
        self.nn_dynamics = nn_dynamics
        self.sparse_dyn = sparse_dyn
        self.batch_size = batch_size
        self.training_epochs = training_epochs
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(process_group) if self.world > 1 else 0
        self.nn_dynamics.to(self.device)
        params = list(self.nn_dynamics.parameters())
        # one flat gradient bucket; every parameter's .grad is a view into it
        self._flat = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=self.device)
        off = 0
        for p in params:
            p.grad = self._flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.optim = torch.optim.Adagrad(params, lr=dyn_lr, weight_decay=weight_decay)
        self.collectives = 0

    # -- moments (mbrl.py:487-509): mean / unbiased std over the data set --------------------------------------------
    def set_moments(self, dataSet):
        col = dataSet.columns()
        x_, x, o_, u = (torch.from_numpy(col[k]) for k in ("x_", "x", "o_", "u"))
        net = self.nn_dynamics
        net.xMean, net.xVar = x_.mean(dim=0, keepdim=True), x_.std(dim=0, keepdim=True)
        dx = x - x_
        if self.sparse_dyn:
            dx = dx[:, dx.shape[1] // 2:]
        net.dxMean, net.dxVar = dx.mean(dim=0, keepdim=True), dx.std(dim=0, keepdim=True)
        net.oMean, net.oVar = o_.mean(dim=0, keepdim=True), o_.std(dim=0, keepdim=True)
        net.uMean, net.uVar = u.mean(dim=0, keepdim=True), u.std(dim=0, keepdim=True)
        net.invalidate()

    def _standardised(self, col):
        """(target dx_norm, network output) for column arrays of transitions (mbrl.py:471-485)."""
        net, dev = self.nn_dynamics, self.device
        t = lambda a: torch.as_tensor(a, dtype=torch.float32, device=dev)
        x_, x, o_, u = t(col["x_"]), t(col["x"]), t(col["o_"]), t(col["u"])
        o_n = (o_ - net.oMean.to(dev)) / net.oVar.to(dev)
        u_n = (u - net.uMean.to(dev)) / net.uVar.to(dev)
        dx = x - x_
        if self.sparse_dyn:
            dx = dx[:, dx.shape[1] // 2:]
        dx_n = (dx - net.dxMean.to(dev)) / net.dxVar.to(dev)
        return dx_n, net(o_n, u_n)

    def training_data(self, batch):
        """batch: list of transition dicts -> (dx_norm, fOutputs_norm), as the reference returns them."""
        keys = ("x_", "u", "x", "o_")
        col = {k: np.asarray([np.asarray(s[k], dtype=np.float64).ravel() for s in batch]).astype(np.float32) for k in keys}
        return self._standardised(col)

    def pred_loss(self, transition):
        dx_n, f = self.training_data([transition])
        return float(torch.linalg.vector_norm((f - dx_n).detach()).cpu())

    # -- one optimiser step on one batch -------------------------------------------------------------------------------
    def train_step(self, col):
        """col: column arrays of ONE full batch (identical on every rank).  Returns the batch's MSE loss."""
        n_total = len(col["x_"])
        local = {k: v[self.rank::self.world] for k, v in col.items()}
        self._flat.zero_()
        if len(local["x_"]):
            dx_n, f = self._standardised(local)
            # nn.MSELoss over the whole batch = sum of squared errors / (rows * outputs); this rank's share of the sum
            loss = ((f - dx_n) ** 2).sum() / float(n_total * dx_n.shape[1])
            loss.backward()
            loss = loss.detach()
        else:
            loss = torch.zeros((), device=self.device)
        if self.world > 1:
            buf = torch.cat([self._flat, loss.reshape(1)])  # gradient bucket + loss: ONE collective per step
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
            self._flat.copy_(buf[:-1])
            loss = buf[-1]
            self.collectives += 1
        self.optim.step()
        return loss

    def training(self, dataSet, printing=False):
        """`MBRL.training`: moments from the data, then `training_epochs` passes over shuffled batches."""
        self.nn_dynamics.eval()
        self.set_moments(dataSet)
        col = dataSet.columns()
        losses = []
        for epoch in range(self.training_epochs):
            running, it = 0.0, 0
            for it, idx in enumerate(dataSet.batch_indices(self.batch_size, shuffle=True)):
                running += float(self.train_step({k: v[idx] for k, v in col.items()}))
            losses.append(running / max(1, it))
            if printing:
                print('NN dynamics training loss:', losses[-1])
        self.nn_dynamics.invalidate()  # the planner's packed device copy is rebuilt from the new weights
        return losses
