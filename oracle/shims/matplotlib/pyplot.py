"""TEST INFRASTRUCTURE ONLY -- no-op `matplotlib.pyplot`."""


class _Nothing:
    """Absorbs every call / attribute (figures, axes, legends)."""

    def __getattr__(self, name):
        return lambda *a, **k: _Nothing()


def subplots(*a, **k):
    return _Nothing(), _Nothing()


def __getattr__(name):
    return lambda *a, **k: None
