"""Algorithm base class (reference: pygent/algorithms/core.py:3-36)."""


class Algorithm(object):
    def __init__(self, environment, agent, t, dt):
        self.t = t
        self.dt = dt
        self.steps = int(t / dt)  # horizon length (core.py:21)
        self.agent = agent
        self.environment = environment
        self.meanCost = []
        self.totalCost = []
        self.episode = 1

    def run_episode(self):
        raise NotImplementedError

    def learning_curve(self):
        raise NotImplementedError

    # Plotting and animation are out of scope; scripts written against the reference call these after a
    # run (ilqr.py:624-644, nmpc.py:105-124), so they are accepted and do nothing.
    def plot(self, *args, **kwargs):
        return None

    def animation(self, *args, **kwargs):
        return None

