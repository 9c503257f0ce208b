"""TEST INFRASTRUCTURE ONLY -- no-op `seaborn` (plots are out of scope; allsteps.py:28-30)."""


def color_palette(*a, **k):
    return [(0.0, 0.0, 0.0)] * int(k.get("n_colors", 10) or 10)


def set_style(*a, **k):
    return None


class _Nothing:
    def __getattr__(self, name):
        return lambda *a, **k: _Nothing()


def heatmap(*a, **k):
    return _Nothing()


def __getattr__(name):
    return lambda *a, **k: None
