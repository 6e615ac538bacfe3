import json
import os


class TensorBoardLogger:
    """Writes scalars as JSON lines (tensorboard is not installed offline) under save_dir/name/version."""

    def __init__(self, save_dir=".", name="default", version="", **_):
        self.save_dir, self.name, self.version = save_dir, name, version
        self.log_dir = os.path.join(save_dir, name, str(version))
        self._fh = None

    def log_metrics(self, metrics, step=None):
        if self._fh is None:
            os.makedirs(self.log_dir, exist_ok=True)
            self._fh = open(os.path.join(self.log_dir, "scalars.jsonl"), "a")
        self._fh.write(json.dumps({"step": step, **{k: float(v) for k, v in metrics.items()}}) + "\n")
        self._fh.flush()

    def log_hyperparams(self, *a, **k):
        pass
