def __getattr__(name):
    def _unavailable(*a, **k):
        raise RuntimeError(f"matplotlib.pyplot.{name}: plotting is not available in this environment (compat stub)")

    return _unavailable
