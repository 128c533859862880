"""Training of the learned dynamics model: the supervised half of the reference's MBRL
(pygent/algorithms/mbrl.py: `set_moments` :487-509, `training_data` :471-485, `training` :450-469, `pred_loss`
:187-206; optimiser `torch.optim.Adagrad(lr=dyn_lr, weight_decay=weight_decay)` :71-73, MSE on standardised velocity
increments, batch 512).

The optimiser step runs as hand-written sm_100a kernels (csrc/pg_train.cu through `pygent_b200.mlp.TrainKernels`: layer 2's
forward, input-gradient and weight-gradient GEMMs on the tcgen05 tensor cores with fp16 hi/lo split operands, fused Adagrad).
Data parallel: every rank holds the same data set and the same network; a batch is cut into `world_size` strided
slices, each rank back-propagates its slice of the SUM of squared errors divided by the full batch size, the gradients
live in ONE flat fp32 bucket (255,002 elements + the loss) that is all-reduced once per step on the compute stream (NCCL;
gloo in the CPU tests), and every rank applies the identical Adagrad update.  This is the only collective of the learned-
dynamics path besides the final cost/argmin gather.
"""
import copy
import os
import pickle
import time

import numpy as np
import torch
import torch.distributed as dist

from .core import Algorithm


class _KernelBackend(object):
    """The CUDA backend of `DynamicsTrainer`: `pygent_b200.mlp.TrainKernels` (csrc/pg_train.cu) on the module's parameters,
    which are re-pointed at ONE flat fp32 device buffer in torch's parameter order so that the kernels update them in place."""

    def __init__(self, net, device, max_rows):
        from .. import mlp
        from ..lib import PygentB200Error
        device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device()) \
            if torch.cuda.is_available() else None
        if device is None or device.type != "cuda" or not torch.cuda.is_available():
            raise PygentB200Error("DynamicsTrainer needs a CUDA device (B200): the training step is hand-written sm_100a "
                                  "kernels and there is no CPU fallback")
        self.net, self.device = net, device
        net.to(device)
        plist = list(net.parameters())
        flat = torch.cat([p.detach().reshape(-1).float() for p in plist]).contiguous()
        off = 0
        for p in plist:
            p.data = flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.k = mlp.TrainKernels(mlp.make_dims(net.xDim, net.uDim, net.xIsAngle), flat, max_rows=max_rows, device=device)
        self.bucket, self.count = self.k.bucket, self.k.count
        self._seen_version = getattr(net, "_weights_version", 0)

    launches = property(lambda self: self.k.launches)

    def _sync_weights(self):
        """The fp16 hi/lo operand images of W2 follow the flat parameter buffer through the Adagrad kernel only: weights
        changed from outside (`load_state_dict`, a hand-set layer) announce themselves through `net.invalidate()`."""
        v = getattr(self.net, "_weights_version", 0)
        if v != self._seen_version:
            self.k.repack()
            self._seen_version = v

    def set_moments(self, *a):
        self.k.set_moments(*a)

    def set_data(self, col):
        self.k.set_data(col)

    def grad(self, idx, n_total):
        self._sync_weights()
        self.k.grad(torch.as_tensor(np.ascontiguousarray(idx, dtype=np.int32), device=self.device), n_total)

    def grad_device(self, idx, n_total):
        self._sync_weights()
        self.k.grad(idx, n_total)

    def enable_p2p(self, group=None):
        self.k.enable_p2p(group)
        self.p2p = True

    def step_p2p(self, idx, n_total, lr, weight_decay, eps=1e-10):
        """Gradient of this rank's slice straight into its NVLink exchange block, then the fused peer-sum + Adagrad kernel."""
        self._sync_weights()
        self.k.grad(torch.as_tensor(np.ascontiguousarray(idx, dtype=np.int32), device=self.device), n_total, bucket=self.k.p2p_bucket())
        return self.k.update_p2p(lr, weight_decay, eps)

    def update(self, lr, weight_decay, eps=1e-10):
        self.k.update(lr, weight_decay, eps)


class DynamicsTrainer(object):
    """`MBRL.set_moments / training_data / training / pred_loss` (mbrl.py:450-509, :187-206) with the optimiser step on the
    device.  `backend` is for tests only (a CPU stand-in that speaks the same protocol drives the host logic on gloo)."""

    def __init__(self, nn_dynamics, dyn_lr=1e-3, weight_decay=1e-3, batch_size=512, training_epochs=60, sparse_dyn=True,
                 device=None, process_group=None, backend=None, exchange=None):
        if not sparse_dyn:
            raise NotImplementedError("the GPU network path implements the sparse_dyn model (velocity half, mbrl.py:54-57)")
        self.nn_dynamics = nn_dynamics
        self.sparse_dyn = sparse_dyn
        self.batch_size = batch_size
        self.training_epochs = training_epochs
        self.dyn_lr, self.weight_decay = float(dyn_lr), float(weight_decay)
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(process_group) if self.world > 1 else 0
        # a rank's slice of a batch never exceeds ceil(batch_size / world) rows (a data set smaller than one batch is ONE batch)
        self.backend = backend if backend is not None else _KernelBackend(nn_dynamics, device, max_rows=batch_size)
        self.device = getattr(self.backend, "device", torch.device("cpu"))
        self.collectives = 0
        self.steps = 0
        self.use_graph = True       # training(): replay each epoch from a CUDA graph (single process, kernel backend)
        self._epoch_graphs = {}
        # gradient exchange of the data-parallel step: "nccl" = one all-reduce of the bucket + the Adagrad kernel; "p2p" = the
        # fused peer-memory kernel (pg_mlp_train_update_p2p): every rank reads the others' buckets over NVLink and updates
        self.exchange = exchange or os.environ.get("PG_TRAIN_EXCHANGE", "nccl")
        if self.world > 1 and self.exchange == "p2p":
            self.backend.enable_p2p(process_group)

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
        self._push_moments()

    def _push_moments(self):
        net = self.nn_dynamics
        g = lambda t: t.detach().cpu().numpy().ravel()
        self.backend.set_moments(g(net.oMean), g(net.oVar), g(net.uMean), g(net.uVar), g(net.dxMean), g(net.dxVar))

    def set_data(self, dataSet_or_col):
        col = dataSet_or_col if isinstance(dataSet_or_col, dict) else dataSet_or_col.columns()
        self.backend.set_data(col)
        self._push_moments()

    def training_data(self, batch):
        """batch: list of transition dicts -> (dx_norm, fOutputs_norm), as the reference returns them (mbrl.py:471-485)."""
        net, dev = self.nn_dynamics, self.device
        keys = ("x_", "u", "x", "o_")
        col = {k: np.asarray([np.asarray(s[k], dtype=np.float64).ravel() for s in batch]).astype(np.float32) for k in keys}
        t = lambda a: torch.as_tensor(a, dtype=torch.float32, device=dev)
        x_, x, o_, u = t(col["x_"]), t(col["x"]), t(col["o_"]), t(col["u"])
        o_n = (o_ - net.oMean.to(dev)) / net.oVar.to(dev)
        u_n = (u - net.uMean.to(dev)) / net.uVar.to(dev)
        dx = x - x_
        if self.sparse_dyn:
            dx = dx[:, dx.shape[1] // 2:]
        return (dx - net.dxMean.to(dev)) / net.dxVar.to(dev), net(o_n, u_n)

    def pred_loss(self, transition):
        """|| f(o_, u) - dx_norm ||_2 of one transition (mbrl.py:187-206), through the training kernels' forward pass:
        the loss slot of the bucket holds SSE / dx_dim."""
        keys = ("x_", "u", "x", "o_")
        self.backend.set_data({k: np.asarray(transition[k], dtype=np.float64).reshape(1, -1).astype(np.float32) for k in keys})
        self._push_moments()
        self.backend.grad(np.zeros(1, dtype=np.int32), 1)
        return float(np.sqrt(float(self.backend.bucket[-1]) * self.nn_dynamics.dxDim))

    # -- a whole epoch as one CUDA graph -------------------------------------------------------------------------------
    def _graph_ok(self):
        return self.world == 1 and isinstance(self.backend, _KernelBackend) and self.use_graph and \
            os.environ.get("PG_TRAIN_GRAPH", "1") != "0"

    def _epoch_from_graph(self, batches):
        """`for idx in batches: train_step(idx)` replayed from a graph keyed by the batch lengths (the last batch of an epoch
        may be shorter): the epoch's shuffled rows go up in one copy into the static index buffer the captured kernels read.
        Same kernels in the same order as the eager loop: identical losses and weights."""
        be, dev = self.backend, self.device
        lens = tuple(len(b) for b in batches)
        key = (lens, float(self.dyn_lr), float(self.weight_decay))  # the step size and decay are baked into the captured launches
        st = self._epoch_graphs.get(key)
        if st is None:
            self._epoch_graphs.clear()  # a growing data set (MBRL adds transitions every episode) changes the key: keep one graph
            be._sync_weights()
            width = max(lens)
            st = {"idx": torch.zeros((len(lens), width), dtype=torch.int32, device=dev),
                  "host": torch.zeros((len(lens), width), dtype=torch.int32).pin_memory(),
                  "run": torch.zeros(len(lens), dtype=torch.float32, device=dev), "graph": torch.cuda.CUDAGraph()}
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                st["graph"].capture_begin(capture_error_mode="thread_local")
                try:
                    for it, n in enumerate(lens):
                        be.k.grad(st["idx"][it, :n], n)
                        be.k.update(self.dyn_lr, self.weight_decay)
                        st["run"][it:it + 1].copy_(be.bucket[-1:])
                finally:
                    st["graph"].capture_end()
            torch.cuda.current_stream(dev).wait_stream(side)
            self._epoch_graphs[key] = st
        be._sync_weights()
        for it, b in enumerate(batches):
            st["host"][it, :len(b)] = torch.from_numpy(np.ascontiguousarray(b, dtype=np.int32))
        st["idx"].copy_(st["host"], non_blocking=True)
        st["graph"].replay()
        self.steps += len(lens)
        return float(st["run"].sum()) / max(1, len(lens) - 1)

    # -- one optimiser step on one batch -------------------------------------------------------------------------------
    def train_step(self, idx):
        """idx: rows of ONE full batch in the uploaded data set (identical on every rank).  Data parallel: every rank
        back-propagates the strided slice idx[rank::world] of the batch with the full batch size in the MSE denominator,
        ONE sum-all-reduce of the flat bucket (255,002 gradient floats + the loss) follows, and every rank applies the
        identical Adagrad update.  Returns the batch's MSE loss (a view of the bucket's last slot: read it before the next step)."""
        idx = np.asarray(idx)
        if self.world > 1 and self.exchange == "p2p":
            self.steps += 1
            return self.backend.step_p2p(idx[self.rank::self.world], len(idx), self.dyn_lr, self.weight_decay)[0]
        self.backend.grad(idx[self.rank::self.world], len(idx))
        if self.world > 1:
            dist.all_reduce(self.backend.bucket, op=dist.ReduceOp.SUM, group=self.group)  # NCCL, on the compute stream
            self.collectives += 1
        self.backend.update(self.dyn_lr, self.weight_decay)
        self.steps += 1
        return self.backend.bucket[-1]

    def training(self, dataSet, printing=False):
        """`MBRL.training`: moments from the data, then `training_epochs` passes over shuffled batches."""
        self.nn_dynamics.eval()
        self.set_moments(dataSet)
        self.set_data(dataSet)
        losses = []
        for epoch in range(self.training_epochs):
            batches = dataSet.batch_indices(self.batch_size, shuffle=True)
            if self._graph_ok():
                # one process, kernel backend: the whole epoch -- every batch's gradient, Adagrad update and loss read-out --
                # is ONE replay of a CUDA graph over a static index buffer (the step is six launches of ~9 us: launch-bound)
                losses.append(self._epoch_from_graph(batches))
                continue
            run = torch.zeros(len(batches), dtype=torch.float32, device=self.device)
            it = 0
            for it, idx in enumerate(batches):
                run[it] = self.train_step(idx)  # stays on the device: one host read per epoch
            losses.append(float(run.sum()) / max(1, it))
            if printing:
                print('NN dynamics training loss:', losses[-1])
        self.nn_dynamics.invalidate()  # the planner's packed device copy is rebuilt from the new weights
        return losses


class MBRL(Algorithm):
    """Model-based RL outer loop (reference: pygent/algorithms/mbrl.py:26-337): random warm-up episodes fill `D_rand`,
    the learned model is (re)trained on `D_rand + D_RL`, episodes are planned on the learned model (iLQR / NMPC on
    `LearnedStateSpaceModel`, tensor-core rollouts and Jacobians) and executed on the real environment, and transitions
    the model predicts badly are added to `D_RL`.  Same constructor arguments and method names as the reference; plotting
    and animation are out of scope.  `precision` selects the planner's network path ("f16x3": tensor cores at the reference's
    fp32 precision; "bf16": lower-precision tensor cores; "fp32": CUDA cores)."""

    def __init__(self, environment, t, dt,
                 test_t=0.,
                 plotInterval=1,
                 nData=int(1e6),
                 path='../results/mbrl/',
                 checkInterval=2,
                 evalPolicyInterval=100,
                 warm_up_episodes=10,
                 dyn_lr=1e-3,
                 batch_size=512,
                 training_epochs=60,
                 data_ratio=1,
                 aggregation_interval=1,
                 fcost=None,
                 horizon=None,
                 use_mpc=False,
                 ilqr_print=False,
                 ilqr_save=False,
                 ilqr_save_interval=100,
                 print_dyn_error=False,
                 weight_decay=1e-3,
                 data_noise=1e-3,
                 prediction_error_bound=1e-3,
                 ou_theta=1,
                 ou_sigma=0.2,
                 maxIters=500,
                 sparse_dyn=True,
                 precision="f16x3",
                 seed=0):
        from ..data import DataSet
        from ..environments import LearnedStateSpaceModel
        from ..nn_models import NNDynamics
        from .nmpc import NMPC
        if not sparse_dyn:
            raise NotImplementedError("the GPU network path implements the sparse_dyn model (velocity half, mbrl.py:54-57)")
        if environment.batched:
            raise NotImplementedError("MBRL learns from one plant; batched planning is available through NMPC / iLQR")
        self.sparse_dyn = sparse_dyn
        xDim, oDim, uDim = environment.xDim, environment.oDim, environment.uDim
        if horizon is None or not use_mpc:
            horizon = t
        self.dt = dt
        self.rng = np.random.RandomState(seed)
        self.nn_dynamics = NNDynamics(xDim, uDim, oDim=oDim, dxDim=xDim // 2, xIsAngle=environment.xIsAngle)
        self.trainer = DynamicsTrainer(self.nn_dynamics, dyn_lr=dyn_lr, weight_decay=weight_decay, batch_size=batch_size,
                                       training_epochs=training_epochs, sparse_dyn=sparse_dyn, device=environment.device)
        self.nn_environment = LearnedStateSpaceModel(self.nn_dynamics, environment._user_cost, environment.x0, uDim, dt,
                                                     precision=precision, device=environment.device)
        self.nn_environment.uMax = environment.uMax
        self.path = path
        for sub in ("", "plots/", "animations/", "data/"):
            os.makedirs(path + sub, exist_ok=True)
        self.nmpc_algorithm = NMPC(environment, self.nn_environment, t, dt, horizon,
                                   init_optim=False, add_noise=True, path=path, fcost=fcost, fastForward=True,
                                   maxIters=maxIters, step_iterations=1, ilqr_print=ilqr_print, ilqr_save=ilqr_save,
                                   tolFun=1e-4, save_interval=ilqr_save_interval, noise_gain=0.005, ou_theta=ou_theta,
                                   ou_sigma=ou_sigma, constrained=True)
        super(MBRL, self).__init__(environment, self.nmpc_algorithm.agent, t, dt)
        self.D_rand = DataSet(nData, seed=seed)
        self.D_RL = DataSet(nData, seed=seed + 1)
        self.test_t = test_t
        self.plotInterval = plotInterval
        self.evalPolicyInterval = evalPolicyInterval
        self.checkInterval = checkInterval
        self.warm_up = warm_up_episodes * int(t / dt)
        self.batch_size = batch_size
        self.training_epochs = training_epochs
        self.data_ratio = data_ratio
        self.aggregation_interval = aggregation_interval
        self.use_mpc = use_mpc
        self.dynamics_first_trained = False
        self.print_dyn_error = print_dyn_error
        self.data_noise = data_noise
        self.prediction_error_bound = prediction_error_bound
        self.training_time, self.optim_time = [], []
        self.expCost, self.episode_steps = [], []

    def ode(self, t, x, u):
        """[x[n/2:], 1/dt * nn_dynamics.ode(x, u)] (mbrl.py:125-130), evaluated by the learned-dynamics library."""
        return self.nn_environment.ode(t, x, u)

    def _transition(self, c):
        env, n = self.environment, self.rng.normal
        return {'x_': env.x_ + n(0, self.data_noise, env.xDim), 'u': np.asarray(self.agent.u, dtype=float) + n(0, self.data_noise, env.uDim),
                'x': env.x + n(0, self.data_noise, env.xDim), 'o_': env.o_ + n(0, self.data_noise, env.oDim),
                'o': env.o + n(0, self.data_noise, env.oDim), 'c': [c], 't': [env.terminated]}

    def _finish_episode(self, cost, i):
        self.meanCost.append(np.mean(cost))
        self.totalCost.append(np.sum(cost))
        self.episode_steps.append(i)
        self.episode += 1

    def run_random_episode(self):
        """Warm-up episode with uniformly random controls (mbrl.py:208-244)."""
        env = self.environment
        env.reset(env.x0)
        self.agent.reset()
        cost, i = [], 0
        for i, _ in enumerate(np.arange(0, self.t, self.dt)):
            u = self.agent.take_random_action(self.dt, ou_noise=False)
            c = env.step(u, self.dt)
            cost.append(c)
            self.D_rand.force_add_sample(self._transition(c))
            if env.terminated:
                break
        self._finish_episode(cost, i)

    def pred_loss(self, transition):
        return self.trainer.pred_loss(transition)

    def run_episode(self, reinit=True):
        """Plan on the learned model, act on the real environment, keep what the model predicts badly (mbrl.py:132-185)."""
        env, agent = self.environment, self.agent
        start = time.time()
        agent.init_optim(env.x0, init=True)
        self.optim_time.append(time.time() - start)
        env.reset(env.x0)
        agent.reset()
        cost, i = [], 0
        for i, _ in enumerate(np.arange(0, self.t + self.test_t, self.dt)):
            u = agent.take_action(self.dt, env.x) if self.use_mpc else agent.take_action_plan_feedback(self.dt, env.x, i)
            c = env.step(u, self.dt)
            cost.append(c)
            tr = self._transition(c)
            if self.pred_loss(tr) > self.prediction_error_bound:
                self.D_RL.force_add_sample(tr)
            if env.terminated:
                break
        self._finish_episode(cost, i)

    def run_controller(self, x0):
        env, agent = self.environment, self.agent
        env.reset(x0)
        agent.reset()
        for i, _ in enumerate(np.arange(0, self.t + self.test_t, self.dt)):
            env.step(agent.take_action(self.dt, env.x), self.dt)
            if env.terminated:
                break

    def train_dynamics(self):
        """mbrl.py:439-448: train on D_rand + data_ratio x D_RL, then the planner sees the new model."""
        start = time.time()
        data = copy.copy(self.D_rand)
        data.data = list(self.D_rand.data)
        for _ in range(self.data_ratio):
            data.data += self.D_RL.data
        losses = self.trainer.training(data)
        self.training_time.append(time.time() - start)
        return losses

    def training(self, dataSet):
        return self.trainer.training(dataSet)

    def run_learning(self, n):
        """mbrl.py:283-310."""
        for steps in range(1, int(n) + 1):
            if len(self.D_rand.data) < max(self.batch_size, self.warm_up):
                self.run_random_episode()
            else:
                if not self.dynamics_first_trained:
                    self.train_dynamics()
                    self.dynamics_first_trained = True
                elif steps % self.aggregation_interval == 0:
                    self.train_dynamics()
                self.run_episode()
                if not self.use_mpc and self.agent.traj_optimizer.saving:
                    self.agent.traj_optimizer.save()
            if self.episode % self.checkInterval == 0:
                self.save()

    def save(self):
        """Checkpoint in the reference's files (mbrl.py:312-337)."""
        d = self.path + 'data/'
        cpu_state = {k: v.detach().cpu() for k, v in self.nn_dynamics.state_dict().items()}
        for name in ('checkpoint' + str(self.episode - 1) + '.pth', 'checkpoint.pth'):
            torch.save({'nn_dynamics': cpu_state}, d + name)
        self.nn_dynamics.save_moments(str(self.episode - 1), self.path)
        self.D_rand.save(d + 'dataSet_D_rand.p')
        self.D_RL.save(d + 'dataSet_D_RL.p')
        self.D_RL.save(d + 'dataSet_D_RL' + str(self.episode - 1) + '.p')
        with open(d + 'learning_curve.p', 'wb') as fh:
            pickle.dump({'totalCost': self.totalCost, 'meanCost': self.meanCost, 'expCost': self.expCost,
                         'episode_steps': self.episode_steps}, fh)
        for name, v in (('training_time.p', self.training_time), ('optim_time.p', self.optim_time)):
            with open(d + name, 'wb') as fh:
                pickle.dump(v, fh)

    def load(self, episode=''):
        """mbrl.py:339-388."""
        d = self.path + 'data/'
        if os.path.isfile(d + 'checkpoint' + episode + '.pth'):
            self.nn_dynamics.load_state_dict(torch.load(d + 'checkpoint' + episode + '.pth')['nn_dynamics'])
        if os.path.isfile(d + 'moments_dict' + episode + '.p'):
            self.nn_dynamics.load_moments(self.path)
        self.nn_dynamics.invalidate()
        if os.path.isfile(d + 'dataSet_D_rand.p'):
            self.D_rand.load(d + 'dataSet_D_rand.p')
        if os.path.isfile(d + 'dataSet_D_RL' + episode + '.p'):
            self.D_RL.load(d + 'dataSet_D_RL' + episode + '.p')
        if os.path.isfile(d + 'learning_curve.p'):
            with open(d + 'learning_curve.p', 'rb') as fh:
                lc = pickle.load(fh)
            self.meanCost, self.totalCost = lc['meanCost'], lc['totalCost']
            self.expCost, self.episode_steps = lc['expCost'], lc['episode_steps']
            self.episode = len(self.meanCost) + 1
