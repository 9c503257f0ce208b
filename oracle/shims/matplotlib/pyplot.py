"""TEST INFRASTRUCTURE ONLY -- no-op `matplotlib.pyplot`."""


def __getattr__(name):
    return lambda *a, **k: None
