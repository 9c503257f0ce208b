"""TEST INFRASTRUCTURE ONLY -- restatement of the slice of `ruptures==1.1.9` that Topsicle calls.

The reference pins `ruptures==1.1.9` (`/root/reference/setup.py:14`,
`requirements.txt:7`) and calls exactly one thing (`allsteps.py:310-311`):

    algo = rpt.Binseg(model="l2").fit(np.array(y))
    result = algo.predict(pen=4, n_bkps=1)

ruptures is a third-party dependency whose source is NOT under /root/reference
and is not installable here (no network, no wheel in /opt/wheelhouse).  This
module restates the published algorithm of that release:

  * `ruptures/detection/binseg.py`  class Binseg: defaults `min_size=2, jump=5`;
    `fit` reshapes a 1-D signal to (n, 1); `predict` runs `sanity_check` and
    raises `BadSegmentationParameters`; `_seg` greedily adds the best split
    while `len(bkps) - 1 < n_bkps` (n_bkps takes precedence over pen/epsilon);
    `single_bkp(start, end)` scans `bkp in range(start, end, jump)` with
    `bkp - start >= min_size and end - bkp >= min_size`,
    `gain = cost(start,end) - cost(start,bkp) - cost(bkp,end)` and keeps
    `max(gain_list)` over `(gain, bkp)` tuples (ties -> larger bkp).
  * `ruptures/costs/costl2.py`  CostL2.error(start, end) =
    `signal[start:end].var(axis=0).sum() * (end - start)`, min_size = 1.
  * `ruptures/utils/utils.py`  sanity_check(n_samples, n_bkps, jump, min_size).

Parity anchor: the reference's own golden `Topsicle_demo/telolengths_all.csv`
(17 breakpoints) is reproduced byte-for-byte through this restatement
(`oracle/make_golden.py` asserts the md5).  Beyond those 17 reads the ruptures
boundary is pinned only by this restatement.
"""
from math import ceil

import numpy as np

__version__ = "1.1.9-restated"


class BadSegmentationParameters(Exception):
    pass


class NotEnoughPoints(Exception):
    pass


class exceptions:  # namespace parity: ruptures.exceptions.*
    BadSegmentationParameters = BadSegmentationParameters
    NotEnoughPoints = NotEnoughPoints


def sanity_check(n_samples, n_bkps, jump, min_size):
    n_adm_bkps = n_samples // jump
    if n_bkps > n_adm_bkps:
        return False
    if n_bkps * ceil(min_size / jump) * jump + min_size > n_samples:
        return False
    return True


class CostL2:
    model = "l2"
    min_size = 1

    def fit(self, signal):
        self.signal = signal.reshape(-1, 1) if signal.ndim == 1 else signal
        return self

    def error(self, start, end):
        if end - start < self.min_size:
            raise NotEnoughPoints
        return self.signal[start:end].var(axis=0).sum() * (end - start)


class Binseg:
    def __init__(self, model="l2", custom_cost=None, min_size=2, jump=5, params=None):
        if model != "l2" or custom_cost is not None:
            raise NotImplementedError("shim restates model='l2' only")
        self.cost = CostL2()
        self.min_size = max(min_size, self.cost.min_size)
        self.jump = jump
        self.n_samples = None
        self.signal = None
        self._cache = {}

    def _single_bkp(self, start, end):
        key = (start, end)
        if key in self._cache:
            return self._cache[key]
        segment_cost = self.cost.error(start, end)
        if np.isinf(segment_cost) and segment_cost < 0:
            res = (None, 0)
        else:
            gain_list = []
            for bkp in range(start, end, self.jump):
                if bkp - start >= self.min_size and end - bkp >= self.min_size:
                    gain = (segment_cost - self.cost.error(start, bkp)
                            - self.cost.error(bkp, end))
                    gain_list.append((gain, bkp))
            if gain_list:
                gain, bkp = max(gain_list)
                res = (bkp, gain)
            else:
                res = (None, 0)
        self._cache[key] = res
        return res

    def _seg(self, n_bkps=None, pen=None, epsilon=None):
        bkps = [self.n_samples]
        stop = False
        while not stop:
            stop = True
            bounds = [0] + bkps
            new_bkps = [self._single_bkp(s, e) for s, e in zip(bounds[:-1], bounds[1:])]
            bkp, gain = max(new_bkps, key=lambda x: x[1])
            if bkp is None:
                break
            if n_bkps is not None:
                if len(bkps) - 1 < n_bkps:
                    stop = False
            elif pen is not None:
                if gain > pen:
                    stop = False
            elif epsilon is not None:
                bounds = [0] + bkps
                error = sum(self.cost.error(s, e) for s, e in zip(bounds[:-1], bounds[1:]))
                if error > epsilon:
                    stop = False
            if not stop:
                bkps.append(bkp)
                bkps.sort()
        bounds = [0] + bkps
        return {(s, e): self.cost.error(s, e) for s, e in zip(bounds[:-1], bounds[1:])}

    def fit(self, signal):
        self.signal = signal.reshape(-1, 1) if signal.ndim == 1 else signal
        self.n_samples = self.signal.shape[0]
        self.cost.fit(signal)
        self._cache = {}
        return self

    def predict(self, n_bkps=None, pen=None, epsilon=None):
        assert any(p is not None for p in (n_bkps, pen, epsilon)), "Give a parameter."
        if not sanity_check(self.cost.signal.shape[0], 0 if n_bkps is None else n_bkps,
                            self.jump, self.min_size):
            raise BadSegmentationParameters
        partition = self._seg(n_bkps=n_bkps, pen=pen, epsilon=epsilon)
        return sorted(e for s, e in partition.keys())

    def fit_predict(self, signal, n_bkps=None, pen=None, epsilon=None):
        return self.fit(signal).predict(n_bkps=n_bkps, pen=pen, epsilon=epsilon)
