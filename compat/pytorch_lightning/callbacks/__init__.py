import os


class Callback:
    pass


class LearningRateMonitor(Callback):
    def __init__(self, logging_interval=None, **_):
        self.logging_interval = logging_interval

    def on_train_epoch_end(self, trainer, model):
        for i, o in enumerate(trainer.optimizers):
            trainer._record(f"lr-{type(o).__name__}", o.param_groups[0]["lr"])


class ModelCheckpoint(Callback):
    def __init__(self, dirpath=None, save_last=False, monitor=None, save_top_k=1, mode="min",
                 filename="{epoch:03d}", **_):
        self.dirpath, self.save_last, self.monitor = dirpath or "checkpoints", save_last, monitor
        self.save_top_k, self.mode, self.filename = save_top_k, mode, filename
        self.best = []

    def on_train_epoch_end(self, trainer, model):
        if self.save_last:
            trainer.save_checkpoint(os.path.join(self.dirpath, "last.ckpt"), model)
        score = trainer.callback_metrics.get(self.monitor) if self.monitor else None
        if score is None or self.save_top_k == 0:
            return
        path = os.path.join(self.dirpath, f"epoch={trainer.current_epoch:03d}-{self.monitor.replace('/', '_')}={float(score):.5f}.ckpt")
        trainer.save_checkpoint(path, model)
        self.best.append((float(score), path))
        self.best.sort(reverse=(self.mode == "max"))
        while self.save_top_k > 0 and len(self.best) > self.save_top_k:
            _, p = self.best.pop()
            if trainer.is_global_zero and os.path.exists(p):
                os.remove(p)
