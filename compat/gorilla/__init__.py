"""Minimal stand-in for gorilla-core 0.2.5.3 (README.md:29 of the reference), which is not installable offline: exactly the
surface the reference's entry points touch — `Config.fromfile` (train.py:50, test.py:53), `utils.set_cuda_visible_devices`
(train.py:58), `parameter_count` (train.py:121), `solver.BaseSolver` / `save_checkpoint` / `load_checkpoint`
(utils/solver.py:19,67; train.py:90; test.py:97).  No hot-path arithmetic lives here."""
import os

import yaml

from . import solver, utils  # noqa: F401


class Config(dict):
    """YAML -> attribute dictionary with `.get`, nested sections as Config objects (cfg.train_dataset.img_size ...)."""

    @staticmethod
    def _wrap(v):
        if isinstance(v, dict):
            return Config({k: Config._wrap(x) for k, x in v.items()})
        if isinstance(v, list):
            return [Config._wrap(x) for x in v]
        return v

    @classmethod
    def fromfile(cls, filename):
        if not os.path.isfile(filename):
            raise FileNotFoundError(filename)
        with open(filename) as f:
            return cls._wrap(yaml.safe_load(f) or {})

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value


def parameter_count(model):
    """{qualified prefix: number of parameters} like fvcore's; train.py:121 only sums the values of the leaf entries, so the
    per-parameter leaves are what is returned."""
    return {name: p.numel() for name, p in model.named_parameters()}
