"""Subset of the pytorch_lightning 1.9 API that PaSCo's scripts and LightningModule touch
(SURVEY.md §7 step 0).  Control plane only: the loop below drives torch.distributed (NCCL,
one process per GPU launched by torchrun) + DistributedDataParallel; it is not a product
component and is kept deliberately small.

Touched surface (reference file:line): LightningModule.save_hyperparameters
(net_panoptic_sparse.py:91), self.log (359-532), lr_schedulers (768), current_epoch (338),
global_step (770), load_from_checkpoint (scripts/eval.py:69-71); Trainer(...) kwargs and
.fit/.test (scripts/train.py:202-239, scripts/eval.py:60-76)."""
from __future__ import annotations

import inspect
import os
from typing import Any, Dict, List, Optional

import torch
import torch.distributed as dist

from . import callbacks, loggers, strategies, plugins  # noqa: F401

__version__ = "1.9.0-shim"


class _HParams(dict):
    __getattr__ = dict.get

    def __setattr__(self, k, v):
        self[k] = v


class LightningModule(torch.nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        self.trainer: Optional["Trainer"] = None
        self._hparams = _HParams()
        self._logged: Dict[str, Any] = {}

    # -- hparams / checkpoints ------------------------------------------------
    def save_hyperparameters(self, *args, ignore=None):
        frame = inspect.currentframe().f_back
        init_args = inspect.getargvalues(frame)
        hp = {n: init_args.locals[n] for n in init_args.args if n != "self"}
        for n in ([ignore] if isinstance(ignore, str) else (ignore or [])):
            hp.pop(n, None)
        self._hparams.update(hp)

    @property
    def hparams(self):
        return self._hparams

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location=None, strict=True, **override):
        ckpt = torch.load(checkpoint_path, map_location=map_location or "cpu", weights_only=False)
        hp = dict(ckpt.get("hyper_parameters", {}))
        hp.update(override)
        model = cls(**hp)
        model.load_state_dict(ckpt["state_dict"], strict=strict)
        return model

    # -- loop state --------------------------------------------------------------
    @property
    def current_epoch(self) -> int:
        return self.trainer.current_epoch if self.trainer else 0

    @property
    def global_step(self) -> int:
        return self.trainer.global_step if self.trainer else 0

    @property
    def global_rank(self) -> int:
        return dist.get_rank() if dist.is_initialized() else 0

    @property
    def logger(self):
        return self.trainer.logger if self.trainer else None

    def lr_schedulers(self):
        s = self.trainer.lr_schedulers if self.trainer else []
        return s[0] if len(s) == 1 else s

    def optimizers(self):
        o = self.trainer.optimizers if self.trainer else []
        return o[0] if len(o) == 1 else o

    def log(self, name, value, on_step=None, on_epoch=None, sync_dist=False, batch_size=None,
            prog_bar=False, logger=True, **_):
        if isinstance(value, torch.Tensor):
            value = value.detach().float()
            if sync_dist and dist.is_initialized() and dist.get_world_size() > 1:
                value = value.clone()
                dist.all_reduce(value)
                value /= dist.get_world_size()
            value = value.item() if value.numel() == 1 else value
        self._logged[name] = value
        if self.trainer is not None:
            self.trainer._record(name, value)

    def log_dict(self, d, **k):
        for n, v in d.items():
            self.log(n, v, **k)

    def print(self, *a, **k):
        if self.global_rank == 0:
            print(*a, **k)

    # hooks (overridden by the model)
    def configure_optimizers(self):
        raise NotImplementedError

    def training_step(self, batch, batch_idx):
        raise NotImplementedError


class LightningDataModule:
    def __init__(self, *a, **k):
        self.trainer = None

    def prepare_data(self):
        pass

    def setup(self, stage=None):
        pass


def seed_everything(seed: int, workers: bool = False):
    import random
    import numpy as np
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    return seed


def _to_device(x, device):
    if isinstance(x, torch.Tensor):
        return x.to(device, non_blocking=True)
    if isinstance(x, dict):
        return {k: _to_device(v, device) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(_to_device(v, device) for v in x)
    return x


class Trainer:
    def __init__(self, accumulate_grad_batches=1, limit_train_batches=1.0, limit_val_batches=1.0,
                 limit_test_batches=1.0, callbacks=None, resume_from_checkpoint=None, max_epochs=1,
                 gradient_clip_val=None, logger=None, check_val_every_n_epoch=1, accelerator="gpu",
                 strategy=None, num_nodes=1, devices=1, sync_batchnorm=False, plugins=None,
                 deterministic=False, log_every_n_steps=50, profiler=None, **_):
        self.accumulate_grad_batches = int(accumulate_grad_batches)
        self.limits = {"train": limit_train_batches, "val": limit_val_batches, "test": limit_test_batches}
        self.callbacks = list(callbacks or [])
        self.resume_from_checkpoint = resume_from_checkpoint
        self.max_epochs = max_epochs
        self.gradient_clip_val = gradient_clip_val
        self.logger = logger
        self.check_val_every_n_epoch = check_val_every_n_epoch
        self.sync_batchnorm = sync_batchnorm
        self.log_every_n_steps = log_every_n_steps
        self.current_epoch = 0
        self.global_step = 0
        self.optimizers: List[torch.optim.Optimizer] = []
        self.lr_schedulers: List[Any] = []
        self.callback_metrics: Dict[str, Any] = {}
        self.world_size = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.device = torch.device(f"cuda:{self.local_rank}" if accelerator == "gpu" and
                                   torch.cuda.is_available() else "cpu")

    # -- helpers ---------------------------------------------------------------
    @property
    def is_global_zero(self):
        return (dist.get_rank() if dist.is_initialized() else 0) == 0

    def _record(self, name, value):
        self.callback_metrics[name] = value
        if self.logger is not None and self.is_global_zero and not isinstance(value, torch.Tensor):
            self.logger.log_metrics({name: value}, step=self.global_step)

    def _limit(self, n, kind):
        lim = self.limits[kind]
        return int(n * lim) if isinstance(lim, float) else min(n, int(lim))

    def _setup(self, model):
        if self.world_size > 1 and not dist.is_initialized():
            dist.init_process_group("nccl" if self.device.type == "cuda" else "gloo")
        if self.device.type == "cuda":
            torch.cuda.set_device(self.device)
        model.trainer = self
        model.to(self.device)
        if self.sync_batchnorm and self.world_size > 1:
            model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
        return model

    def _configure(self, model):
        cfg = model.configure_optimizers()
        if isinstance(cfg, tuple) and len(cfg) == 2:
            opts, scheds = cfg
            self.optimizers = list(opts)
            self.lr_schedulers = [s["scheduler"] if isinstance(s, dict) else s for s in scheds]
            self._sched_interval = [s.get("interval", "epoch") if isinstance(s, dict) else "epoch" for s in scheds]
        elif isinstance(cfg, dict):
            self.optimizers = [cfg["optimizer"]]
            s = cfg.get("lr_scheduler")
            self.lr_schedulers = [s["scheduler"] if isinstance(s, dict) else s] if s else []
            self._sched_interval = [s.get("interval", "epoch") if isinstance(s, dict) else "epoch"] if s else []
        else:
            self.optimizers = [cfg]

    def save_checkpoint(self, path, model):
        if not self.is_global_zero:
            return
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        torch.save({"state_dict": model.state_dict(), "hyper_parameters": dict(model.hparams),
                    "epoch": self.current_epoch, "global_step": self.global_step,
                    "optimizer_states": [o.state_dict() for o in self.optimizers],
                    "lr_schedulers": [s.state_dict() for s in self.lr_schedulers]}, path)

    # -- loops -------------------------------------------------------------------
    def fit(self, model, datamodule=None, train_dataloaders=None, val_dataloaders=None, ckpt_path=None):
        model = self._setup(model)
        if datamodule is not None:
            datamodule.setup("fit")
            train_dataloaders = datamodule.train_dataloader()
            val_dataloaders = datamodule.val_dataloader()
        self._configure(model)
        resume = ckpt_path or self.resume_from_checkpoint
        if resume:
            ck = torch.load(resume, map_location="cpu", weights_only=False)
            model.load_state_dict(ck["state_dict"])
            for o, s in zip(self.optimizers, ck.get("optimizer_states", [])):
                o.load_state_dict(s)
            for sc, s in zip(self.lr_schedulers, ck.get("lr_schedulers", [])):
                sc.load_state_dict(s)
            self.current_epoch, self.global_step = ck.get("epoch", 0) + 1, ck.get("global_step", 0)
        opt = self.optimizers[0]
        # world > 1: the reference's DDP(find_unused_parameters=True) (scripts/train.py:213) = pasco_b200's bucketed reducer:
        # gradients are views into one flat buffer, buckets are all-reduced from backward hooks while backward still runs
        reducer = None
        if self.world_size > 1:
            from pasco_b200.parallel import GradReducer
            reducer = GradReducer(list(model.parameters()))
        A = self.accumulate_grad_batches
        for epoch in range(self.current_epoch, self.max_epochs):
            self.current_epoch = epoch
            model.train()
            n = self._limit(len(train_dataloaders), "train")
            for i, batch in enumerate(train_dataloaders):
                if i >= n:
                    break
                batch = _to_device(batch, self.device)
                last = (i + 1) % A == 0
                if reducer is not None:
                    reducer.rearm(sync=last)        # only the last micro-step of an optimiser step launches the all-reduces
                out = model.training_step(batch, i)
                loss = out["loss"] if isinstance(out, dict) else out
                (loss / A).backward()
                if last:
                    if reducer is not None:
                        reducer.finish()
                    if self.gradient_clip_val:
                        torch.nn.utils.clip_grad_norm_(model.parameters(), self.gradient_clip_val)
                    opt.step()
                    if reducer is not None:
                        reducer.zero_grad()         # one memset; the gradients stay views into the flat buffer
                    else:
                        opt.zero_grad(set_to_none=True)
                    self.global_step += 1
            for sc, iv in zip(self.lr_schedulers, getattr(self, "_sched_interval", [])):
                if iv == "epoch":                   # Lightning steps an {"interval": "epoch"} scheduler once per epoch
                    sc.step()                       # (net_panoptic_sparse.py:901; the model also steps it per batch, :768-770)
            if val_dataloaders is not None and (epoch + 1) % self.check_val_every_n_epoch == 0:
                self._eval_loop(model, val_dataloaders, "val")
            for cb in self.callbacks:
                if hasattr(cb, "on_train_epoch_end"):
                    cb.on_train_epoch_end(self, model)
        if reducer is not None:
            reducer.detach()
        return model

    def _eval_loop(self, model, loader, kind):
        model.eval()
        outs = []
        step = getattr(model, f"{'validation' if kind == 'val' else 'test'}_step")
        with torch.no_grad():
            n = self._limit(len(loader), kind)
            for i, batch in enumerate(loader):
                if i >= n:
                    break
                outs.append(step(_to_device(batch, self.device), i))
        end = getattr(model, f"{'validation' if kind == 'val' else 'test'}_epoch_end", None)
        if end is not None:
            end(outs)
        return outs

    def validate(self, model=None, dataloaders=None, datamodule=None, **_):
        model = self._setup(model)
        if datamodule is not None:
            datamodule.setup("validate")
            dataloaders = datamodule.val_dataloader()
        return self._eval_loop(model, dataloaders, "val")

    def test(self, model=None, dataloaders=None, datamodule=None, **_):
        model = self._setup(model)
        if datamodule is not None:
            datamodule.setup("test")
            dataloaders = datamodule.test_dataloader()
        return self._eval_loop(model, dataloaders, "test")
