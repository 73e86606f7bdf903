"""Stand-in for tensorboardX.SummaryWriter (train.py:30,71,203): scalars are appended to <log_dir>/scalars.tsv."""
import os


class SummaryWriter:
    def __init__(self, log_dir=None, **kwargs):
        self.log_dir = log_dir or "runs"
        os.makedirs(self.log_dir, exist_ok=True)
        self._f = open(os.path.join(self.log_dir, "scalars.tsv"), "a")

    def add_scalar(self, tag, value, global_step=None, **kwargs):
        self._f.write(f"{tag}\t{global_step}\t{float(value)}\n")
        self._f.flush()

    def close(self):
        self._f.close()
