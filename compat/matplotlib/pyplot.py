"""Null plotter: every pyplot call returns an object that accepts any further call and draws nothing, so that the reference's mAP
evaluation (utils/evaluation_utils.py:877-949 builds its figures unconditionally) runs to its numbers in an environment without
matplotlib.  Nothing is written to disk."""


class _Null:
    def __call__(self, *a, **k):
        return self

    def __getattr__(self, name):
        return self

    def __iter__(self):
        return iter(())


_NULL = _Null()


def __getattr__(name):
    return _NULL
