"""Control history holders (reference: pygent/agents.py:11-113), batched.

`history` has the reference's layout for a single plant (`[k+1, uDim]`, row 0 is the zero dummy the
reference keeps, `uu = history[1:]`, ilqr.py:227) and `[B, k+1, uDim]` for a batch.
"""
import numpy as np


class Agent(object):
    def __init__(self, uDim, batch=None):
        self.uDim = uDim
        self.batch = batch
        self.reset()

    def take_action(self, *args):
        raise NotImplementedError

    def control(self, dt, u):
        """Apply (record) the given control; returns it (agents.py:30-46)."""
        self.u = u
        self._hist.append(np.asarray(u, dtype=float).reshape(self._hist[0].shape))
        self.tt.extend([self.tt[-1] + dt])
        return self.u

    def reset(self):
        self.tt = [0]
        shape = (self.uDim,) if self.batch is None else (self.batch, self.uDim)
        self._hist = [np.zeros(shape)]

    @property
    def history(self):
        return np.stack(self._hist, axis=0 if self.batch is None else 1)

    def set_history(self, uu):
        """uu: [T, m] or [B, T, m] -> history with the zero dummy row in front."""
        uu = np.asarray(uu, dtype=float)
        if self.batch is None:
            self._hist = [np.zeros(self.uDim)] + list(uu)
        else:
            self._hist = [np.zeros((self.batch, self.uDim))] + [uu[:, k] for k in range(uu.shape[1])]
        self.tt = [0]

    def plot(self, *args, **kwargs):  # out of scope; accepted, does nothing (agents.py:60-84)
        return None

    def save_history(self, filename, path):
        import pickle
        with open(path + filename + ".p", "wb") as fh:
            pickle.dump({"tt": self.tt, "uu": self.history}, fh)


class FeedBack(Agent):
    """State feedback u = mu(x) (agents.py:85-113)."""

    def __init__(self, mu, uDim, batch=None):
        super(FeedBack, self).__init__(uDim, batch)
        self.mu = mu

    def take_action(self, dt, x):
        return self.control(dt, self.mu(x))
