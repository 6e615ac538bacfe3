"""CPU tests of the pytorch_lightning import shim (`compat/pytorch_lightning`) with a toy module that touches exactly the
Lightning surface the reference touches (net_panoptic_sparse.py:91 save_hyperparameters, :359 self.log, :768-770
lr_schedulers().step(global_step), :779 validation_epoch_end, :887-903 configure_optimizers → ([opt], [{"scheduler",
"interval"}]); scripts/train.py:181-239 Trainer kwargs, ModelCheckpoint, LearningRateMonitor, TensorBoardLogger,
DDPStrategy, SLURMEnvironment, resume_from_checkpoint; scripts/eval.py:60-76 load_from_checkpoint + trainer.test).
World 2 over gloo: the shim's training loop with the bucketed reducer equals one process stepping on the global batch."""
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _pl():
    for p in (ROOT, os.path.join(ROOT, "compat")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import pytorch_lightning as pl
    assert pl.__version__.endswith("shim")
    return pl


def _toy(pl):
    class Toy(pl.LightningModule):
        def __init__(self, width=16, lr=1e-2, unused=True):
            super().__init__()
            self.save_hyperparameters()
            self.body = torch.nn.Sequential(torch.nn.Linear(8, width), torch.nn.ReLU(), torch.nn.Linear(width, 4))
            self.spare = torch.nn.Linear(4, 4) if unused else None      # never used in forward: find_unused_parameters
            self.lr = lr
            self.seen = {"val": 0, "test": 0, "val_end": 0, "test_end": 0}

        def training_step(self, batch, batch_idx):
            self.lr_schedulers().step(self.global_step)
            x, y = batch
            loss = (self.body(x) - y).square().mean()
            self.log("train/loss", loss.detach(), on_epoch=True, sync_dist=True)
            return {"loss": loss}

        def validation_step(self, batch, batch_idx):
            self.seen["val"] += 1
            x, y = batch
            return {"l": (self.body(x) - y).square().mean()}

        def test_step(self, batch, batch_idx):
            self.seen["test"] += 1
            return self.validation_step(batch, batch_idx)

        def validation_epoch_end(self, outputs):
            self.seen["val_end"] += 1
            self.log("val/loss", torch.stack([o["l"] for o in outputs]).mean(), sync_dist=True)

        def test_epoch_end(self, outputs):
            self.seen["test_end"] += 1

        def configure_optimizers(self):
            opt = torch.optim.AdamW(self.parameters(), lr=self.lr, weight_decay=0.0)
            sch = torch.optim.lr_scheduler.LambdaLR(opt, lambda step: 1.0 / (1.0 + 0.1 * step))
            return [opt], [{"scheduler": sch, "interval": "epoch"}]
    return Toy


def _batches(n, seed, bs=4):
    g = torch.Generator().manual_seed(seed)
    return [(torch.randn(bs, 8, generator=g), torch.randn(bs, 4, generator=g)) for _ in range(n)]


def test_fit_checkpoint_resume_and_eval(tmp_path):
    pl = _pl()
    from pytorch_lightning.callbacks import LearningRateMonitor, ModelCheckpoint
    from pytorch_lightning.loggers import TensorBoardLogger
    from pytorch_lightning.plugins.environments import SLURMEnvironment
    from pytorch_lightning.strategies import DDPStrategy
    Toy = _toy(pl)
    pl.seed_everything(3)
    model = Toy(width=12, lr=5e-3)
    assert dict(model.hparams) == {"width": 12, "lr": 5e-3, "unused": True}
    train, val = _batches(8, 1), _batches(3, 2)
    ck = ModelCheckpoint(dirpath=str(tmp_path / "ck"), save_last=True, monitor="val/loss", save_top_k=1, mode="min")
    logger = TensorBoardLogger(save_dir=str(tmp_path / "tb"), name="toy", version="")
    kw = dict(accumulate_grad_batches=2, limit_train_batches=1.0, limit_val_batches=1.0, gradient_clip_val=0.5, logger=logger,
              check_val_every_n_epoch=1, accelerator="cpu", strategy=DDPStrategy(find_unused_parameters=True), num_nodes=1,
              devices=1, sync_batchnorm=True, plugins=[SLURMEnvironment(requeue_signal=None)])
    tr = pl.Trainer(callbacks=[ck, LearningRateMonitor(logging_interval="step")], max_epochs=2, **kw)
    w0 = model.body[0].weight.detach().clone()
    tr.fit(model, train_dataloaders=train, val_dataloaders=val)
    assert tr.global_step == 8 and tr.current_epoch == 1              # 8 batches / accum 2 = 4 optimiser steps per epoch
    assert model.seen["val"] == 6 and model.seen["val_end"] == 2
    assert not torch.equal(w0, model.body[0].weight)
    assert "val/loss" in tr.callback_metrics and "train/loss" in tr.callback_metrics and "lr-AdamW" in tr.callback_metrics
    last = os.path.join(str(tmp_path / "ck"), "last.ckpt")
    assert os.path.exists(last) and len([f for f in os.listdir(str(tmp_path / "ck")) if f.startswith("epoch=")]) == 1
    lines = [json.loads(l) for l in open(os.path.join(logger.log_dir, "scalars.jsonl"))]
    assert any("val/loss" in l for l in lines)
    # eval.py path: load_from_checkpoint rebuilds the module from the saved hyper-parameters, trainer.test runs the test hooks
    again = Toy.load_from_checkpoint(checkpoint_path=last)
    assert again.hparams.width == 12
    for a, b in zip(again.state_dict().values(), model.state_dict().values()):
        assert torch.equal(a, b)
    outs = pl.Trainer(accelerator="cpu", devices=1, logger=False).test(model=again, dataloaders=val)
    assert len(outs) == 3 and again.seen["test"] == 3 and again.seen["test_end"] == 1
    # train.py resume path: a third epoch from last.ckpt continues epoch / step counters and the optimiser state
    tr2 = pl.Trainer(callbacks=[], max_epochs=3, resume_from_checkpoint=last, **kw)
    cont = Toy(width=12, lr=5e-3)
    tr2.fit(cont, train_dataloaders=train, val_dataloaders=val)
    assert tr2.current_epoch == 2 and tr2.global_step == 12
    straight = Toy(width=12, lr=5e-3)
    pl.seed_everything(3)
    straight = Toy(width=12, lr=5e-3)
    pl.Trainer(callbacks=[], max_epochs=3, **kw).fit(straight, train_dataloaders=train, val_dataloaders=val)
    for a, b in zip(cont.state_dict().values(), straight.state_dict().values()):
        assert torch.allclose(a, b, atol=1e-6), "resumed run differs from the uninterrupted one"


def _ddp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world), RANK=str(rank),
                      LOCAL_RANK=str(rank))
    pl = _pl()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    Toy = _toy(pl)
    torch.manual_seed(11)
    model = Toy(width=16, lr=1e-2)
    tr = pl.Trainer(accumulate_grad_batches=2, max_epochs=2, accelerator="cpu", gradient_clip_val=0.5, devices=world)
    assert tr.world_size == world
    tr.fit(model, train_dataloaders=_batches(4, 50 + rank), val_dataloaders=_batches(1, 60 + rank))
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    gathered = [None] * world
    dist.all_gather_object(gathered, sd)
    ok = all(torch.equal(gathered[0][k], gathered[r][k]) for r in range(world) for k in sd)       # replicas stay in lock-step
    ok &= abs(tr.callback_metrics["train/loss"] - model._logged["train/loss"]) < 1e-12
    if rank == 0:
        # one process on the global batch: the mean over ranks of the per-rank mean losses
        torch.manual_seed(11)
        ref = Toy(width=16, lr=1e-2)
        opt = torch.optim.AdamW(ref.parameters(), lr=1e-2, weight_decay=0.0)
        data = [_batches(4, 50 + r) for r in range(world)]
        step = 0
        for _ in range(2):
            for i in range(4):
                for g in opt.param_groups:
                    g["lr"] = 1e-2 / (1.0 + 0.1 * step)
                loss = sum((ref.body(x) - y).square().mean() for x, y in (d[i] for d in data)) / world
                (loss / 2).backward()
                if i % 2 == 1:
                    torch.nn.utils.clip_grad_norm_([p for p in ref.parameters() if p.grad is not None], 0.5)
                    opt.step()
                    opt.zero_grad(set_to_none=True)
                    step += 1
        ok &= all(torch.allclose(sd[k], v, atol=1e-6) for k, v in ref.state_dict().items())
        ok &= step == tr.global_step == 4
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_world2_fit_equals_single_process_on_the_global_batch():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_ddp_worker, args=(world, 29500 + os.getpid() % 400, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}, dict(out)
