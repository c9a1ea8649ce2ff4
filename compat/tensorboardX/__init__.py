"""Stand-in for tensorboardX (utils/solver.py:12, not installable offline): scalars go to a CSV-like text file in the log directory."""
import os


class SummaryWriter:
    def __init__(self, logdir=None, **kwargs):
        self.logdir = logdir or "."
        os.makedirs(self.logdir, exist_ok=True)
        self._f = open(os.path.join(self.logdir, "scalars.txt"), "a")

    def add_scalar(self, tag, value, global_step=None, **kwargs):
        self._f.write(f"{tag}\t{global_step}\t{float(value)}\n")

    def flush(self):
        self._f.flush()

    def close(self):
        self._f.close()
