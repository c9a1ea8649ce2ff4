"""gorilla.solver: BaseSolver + checkpoint I/O as used by utils/solver.py and train.py / test.py of the reference."""
import os
from collections import OrderedDict

import torch


class LogBuffer:
    """History of scalar dictionaries: update() appends, average(n) fills `_output` with the mean of the last n entries (all when
    n == 0), `avg` is the mean over the whole history (utils/solver.py:107-125)."""

    def __init__(self):
        self.val_history = OrderedDict()
        self._output = OrderedDict()

    def clear(self):
        self.val_history.clear()
        self._output.clear()

    def update(self, variables, count=1):
        for k, v in variables.items():
            self.val_history.setdefault(k, []).append(float(v))

    def average(self, n=0):
        for k, vals in self.val_history.items():
            tail = vals[-n:] if n > 0 else vals
            self._output[k] = sum(tail) / max(len(tail), 1)

    @property
    def output(self):
        return self._output

    @property
    def avg(self):
        return OrderedDict((k, sum(v) / max(len(v), 1)) for k, v in self.val_history.items())


class _NullWriter:
    def add_scalar(self, *a, **k):
        pass


class BaseSolver:
    def __init__(self, model, dataloaders, cfg, logger=None, **kwargs):
        self.model = model
        self.dataloaders = dataloaders
        self.cfg = cfg
        self.logger = logger
        self.log_buffer = LogBuffer()
        self.tb_writer = _NullWriter()
        self.epoch = kwargs.get("start_epoch", 1)
        self.iter = kwargs.get("start_iter", 0)


def _unwrap(model):
    return model.module if hasattr(model, "module") and isinstance(model, torch.nn.DataParallel) else model


def save_checkpoint(model, filename, optimizer=None, scheduler=None, meta=None):
    """File layout the reference reads back: {'model', 'optimizer', 'meta'} (train.py:91-92,107)."""
    os.makedirs(os.path.dirname(os.path.abspath(filename)), exist_ok=True)
    ckpt = {"meta": dict(meta or {}), "model": OrderedDict((k, v.detach().cpu()) for k, v in _unwrap(model).state_dict().items())}
    if optimizer is not None:
        ckpt["optimizer"] = optimizer.state_dict()
    if scheduler is not None:
        ckpt["scheduler"] = scheduler.state_dict()
    torch.save(ckpt, filename)


def load_checkpoint(model, filename, map_location="cpu", strict=True, optimizer=None, **kwargs):
    ckpt = torch.load(filename, map_location=map_location, weights_only=False)
    state = ckpt["model"] if "model" in ckpt else ckpt
    state = OrderedDict((k[7:] if k.startswith("module.") else k, v) for k, v in state.items())
    _unwrap(model).load_state_dict(state, strict=strict)
    if optimizer is not None and "optimizer" in ckpt:
        optimizer.load_state_dict(ckpt["optimizer"])
    return ckpt
