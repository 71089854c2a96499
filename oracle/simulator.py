"""
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  CPU restatement of the discrete-time half of zopt/simulator.py:
two blocks `(y0, x0') = blocks[0].update(k, x0, x1)`, `(y1, x1') = blocks[1].update(k, x1, y0)` stepped N = ceil(T/dt)
times (simulator.py:124-138), outputs re-evaluated along the stored states (simulator.py:165-168).  Blocks are plain
callables on NumPy arrays, exactly like the reference's lambdas.
"""
import numpy as np


class SimBlock:  # zopt/simulator.py:9-39 (no jit here)
    def __init__(self, fun, x0, dt=0, jittable=True, name=None):
        self.update, self.dt, self.jittable, self.x0, self.nx, self.name = fun, dt, jittable, np.asarray(x0, dtype=float), len(x0), name


class Simulator:  # zopt/simulator.py:48-169, discrete branch
    def __init__(self, blocks, t_span, method="RK45", t_eval=None):
        assert len(blocks) == 2 and len({b.dt for b in blocks}) == 1 and blocks[0].dt > 0
        self.blocks, self.t_span, self.dt = blocks, t_span, blocks[0].dt

    def _step(self, k, x):  # simulator.py:124-129
        n0 = self.blocks[0].nx
        x0, x1 = x[:n0], x[n0:]
        y0, x0 = self.blocks[0].update(k, x0, x1)
        y1, x1 = self.blocks[1].update(k, x1, y0)
        return np.concatenate([np.asarray(x0, dtype=float).reshape(-1), np.asarray(x1, dtype=float).reshape(-1)])

    def simulate(self):  # simulator.py:131-169
        x0 = np.concatenate([b.x0 for b in self.blocks])
        N = int(np.ceil(self.t_span[1] / self.dt))
        xArr = np.zeros((N + 1, len(x0)))
        xArr[0] = x0
        for k in range(N):
            xArr[k + 1] = self._step(k, xArr[k])
        tArr = np.arange(0, N + 1) * self.dt
        kArr = np.arange(0, len(tArr) - 1)
        n0 = self.blocks[0].nx
        x0Arr, x1Arr = xArr[:, :n0], xArr[:, n0:]
        y0Arr = np.array([self.blocks[0].update(k, a, b)[0] for (k, a, b) in zip(kArr, x0Arr, x1Arr)])
        y1Arr = np.array([self.blocks[1].update(k, b, y)[0] for (k, b, y) in zip(kArr, x1Arr, y0Arr)])
        return tArr, x0Arr, x1Arr, y0Arr, y1Arr
