"""Host-side helpers of the hot path (reference: pygent/helpers.py).

Only what the simulation-and-planning path touches: `observation` (:20-29), `mapAngles` (:204-213)
and the central-difference formulas (:41-148, eps = 1e-3).  The device kernels carry their own
copies of `observation` (pg_observe) and `system_linearization` (lin_mode 1); these numpy versions
exist for API compatibility and for checking the kernels.
"""
import numpy as np


def observation(x, xIsAngle):
    """Angles become [cos, sin] (in that order), other states pass through (helpers.py:20-29).
    Accepts `[n]` or `[B, n]`."""
    x = np.asarray(x, dtype=float)
    cols = []
    for i, is_angle in enumerate(xIsAngle):
        if is_angle:
            cols.append(np.cos(x[..., i]))
            cols.append(np.sin(x[..., i]))
        else:
            cols.append(x[..., i])
    return np.stack(cols, axis=-1)


def mapAngles(xIsAngle, x, mod=np):
    """Maps angles to the interval [-pi, pi) (helpers.py:204-213)."""
    return [(s + mod.pi) % (2 * mod.pi) - mod.pi if a else s for a, s in zip(xIsAngle, x)]


def _unit(n, i, eps):
    e = np.zeros(n)
    e[i] = eps
    return e


def system_linearization(f, xx, uu, eps=1e-3):
    """A = df/dx, B = df/du by central differences (helpers.py:130-148)."""
    xx, uu = np.asarray(xx, dtype=float), np.asarray(uu, dtype=float)
    cols_a = [(np.asarray(f(xx + _unit(len(xx), i, eps), uu)) - np.asarray(f(xx - _unit(len(xx), i, eps), uu))) / (2 * eps)
              for i in range(len(xx))]
    cols_b = [(np.asarray(f(xx, uu + _unit(len(uu), j, eps))) - np.asarray(f(xx, uu - _unit(len(uu), j, eps)))) / (2 * eps)
              for j in range(len(uu))]
    return np.stack(cols_a, axis=1), np.stack(cols_b, axis=1)


def fx(f, xx, uu, eps=1e-3):
    xx = np.asarray(xx, dtype=float)
    return np.stack([np.atleast_1d((f(xx + _unit(len(xx), i, eps), uu) - f(xx - _unit(len(xx), i, eps), uu)) / (2 * eps))
                     for i in range(len(xx))], axis=1)


def fu(f, xx, uu, eps=1e-3):
    uu = np.asarray(uu, dtype=float)
    return np.stack([np.atleast_1d((f(xx, uu + _unit(len(uu), j, eps)) - f(xx, uu - _unit(len(uu), j, eps))) / (2 * eps))
                     for j in range(len(uu))], axis=1)


def fxx(f, xx, uu, eps=1e-3):
    xx = np.asarray(xx, dtype=float)
    n = len(xx)
    H = np.zeros((n, n))
    for i in range(n):
        for j in range(n):
            ei, ej = _unit(n, i, eps), _unit(n, j, eps)
            H[i, j] = (f(xx + ei + ej, uu) - f(xx - ei + ej, uu) - f(xx + ei - ej, uu) + f(xx - ei - ej, uu)) / (4 * eps ** 2)
    return H


def fuu(f, xx, uu, eps=1e-3):
    uu = np.asarray(uu, dtype=float)
    m = len(uu)
    H = np.zeros((m, m))
    for i in range(m):
        for j in range(m):
            ei, ej = _unit(m, i, eps), _unit(m, j, eps)
            H[i, j] = (f(xx, uu + ei + ej) - f(xx, uu - ei + ej) - f(xx, uu + ei - ej) + f(xx, uu - ei - ej)) / (4 * eps ** 2)
    return H


def fxu(f, xx, uu, eps=1e-3):
    xx, uu = np.asarray(xx, dtype=float), np.asarray(uu, dtype=float)
    n, m = len(xx), len(uu)
    H = np.zeros((n, m))
    for i in range(n):
        for j in range(m):
            ex, eu = _unit(n, i, eps), _unit(m, j, eps)
            H[i, j] = (f(xx + ex, uu + eu) - f(xx - ex, uu + eu) - f(xx + ex, uu - eu) + f(xx - ex, uu - eu)) / (4 * eps ** 2)
    return H


class OUnoise(object):
    """Ornstein-Uhlenbeck exploration noise, discrete Euler form (reference: pygent/helpers.py:170-203):
    dx = theta (mu - x) dt + sigma sqrt(dt) N(0, 1).  `shape` may carry a batch axis."""

    def __init__(self, uDim, dt, mu=0, theta=0.15, sigma=0.2, seed=0, batch=None):
        self.uDim, self.mu, self.theta, self.sigma, self.dt = uDim, mu, theta, sigma, dt
        self.shape = (uDim,) if batch is None else (batch, uDim)
        self.rng = np.random.RandomState(seed)
        self.reset()

    def reset(self):
        self.x = np.ones(self.shape) * self.mu

    def sample(self):
        dx = self.theta * (self.mu - self.x) * self.dt + self.sigma * np.sqrt(self.dt) * self.rng.normal(size=self.shape)
        self.x = self.x + dx
        return self.x
