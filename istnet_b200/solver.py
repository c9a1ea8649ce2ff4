"""Training loop with the interface and schedule of the reference's `Solver` (utils/solver.py:19-215), re-built around the captured
step ("remove the per-step cuda.synchronize + 3x .item() stalls and fuse Adam + CyclicLR + BN-momentum update").

What the reference does every iteration, and what replaces it here:

  utils/solver.py:153       torch.cuda.synchronize()                 -> nothing: the loop never waits for the device
  :158-161                  .cuda() of every tensor of both batches  -> one pinned staging set, asynchronous copies on a copy stream
                                                                         into a device staging set, device-to-device into the graph's
                                                                         static inputs (the copy of batch i+1 overlaps step i)
  :163-174                  10 torch.cat                             -> syn | real are written side by side into the staging set
  :175-182                  model forward, 2 x SupervisedLoss        -> inside the CUDA graph (graph.GraphedTrainStep)
  :184-186                  3 x .item()                              -> the three losses leave the graph as one 3-float tensor, copied to a
                                                                         pinned ring; the host reads slot i - log_lag (already complete)
  :88-92                    lr_scheduler.step / bnm_scheduler.step   -> closed forms on the host (cyclic_lr, bn_momentum_at) written to DEVICE
                                                                         scalars the captured kernels read (FlatAdam.lr_dev, nhwc momentum table)
  :94-99                    zero_grad / backward / Adam.step         -> inside the graph (flat gradient buckets, csrc/optim.cu)

`log_lag=0` reproduces the reference's per-iteration blocking reads (loss of iteration i logged at iteration i); the default 1 logs the
loss of iteration i when iteration i+1 has been enqueued — same numbers, one iteration later, no stall.

The object keeps the reference's attribute names (model, dataloaders, cfg, logger, log_buffer, tb_writer, optimizer, epoch, iter,
per_write) and methods (solve / train / get_logger_info / write_summary), so train.py:136-147 can construct it in place of
`Solver`.  CUDA only: like the rest of the package there is no CPU fallback."""
import os
import time
from collections import OrderedDict

import torch

MODEL_IN = ("rgb", "pts", "choose", "category_label", "qo")
LABELS = ("qo", "rotation_label", "translation_label", "size_label")


def cyclic_lr(it, base_lr=1e-5, max_lr=1e-3, step_size_up=1):
    """torch.optim.lr_scheduler.CyclicLR(mode='triangular', cycle_momentum=False) evaluated at iteration `it`
    (utils/solver.py:46-47 builds it with step_size_up = max_epoch * num_mini_batch_per_epoch // 6 and calls step(self.iter))."""
    total = 2.0 * step_size_up
    cycle = int(1 + it / total)
    x = 1.0 + it / total - cycle
    scale = x / 0.5 if x <= 0.5 else (x - 1.0) / (0.5 - 1.0)
    return base_lr + (max_lr - base_lr) * scale


def bn_momentum_at(it, bn_momentum, bn_decay, decay_step, bnm_clip):
    """utils/solver.py:49: max(bn_momentum * bn_decay ** int(it / decay_step), bnm_clip)."""
    return max(bn_momentum * bn_decay ** int(it / decay_step), bnm_clip)


class LogBuffer:
    """gorilla's LogBuffer as the reference uses it (utils/solver.py:107-125): update / average(n) -> _output / avg / clear."""

    def __init__(self):
        self.val_history, self._output = OrderedDict(), OrderedDict()

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
    def avg(self):
        return OrderedDict((k, sum(v) / max(len(v), 1)) for k, v in self.val_history.items())


def merge_into(dst, syn, real, keys):
    """utils/solver.py:163-174 without the temporaries: dst[k][:b1] = syn[k], dst[k][b1:] = real[k]."""
    b1 = syn[keys[0]].shape[0]
    for k in keys:
        dst[k][:b1].copy_(syn[k].reshape(dst[k][:b1].shape))
        dst[k][b1:].copy_(real[k].reshape(dst[k][b1:].shape))
    return b1


class Solver:
    RING = 8  # pinned loss slots / maximum number of iterations the host may run ahead of the device

    def __init__(self, model, data_mode, loss, dataloaders, logger, cfg, start_epoch=1, start_iter=0, tb_writer=None, log_lag=1,
                 process_group=None):
        assert torch.cuda.is_available(), "istnet_b200.solver.Solver needs a GPU; there is no CPU fallback"
        self.model, self.data_mode, self.loss, self.dataloaders, self.logger, self.cfg = model, data_mode, loss, dataloaders, logger, cfg
        self.log_buffer, self.tb_writer = LogBuffer(), tb_writer
        self.per_val, self.per_write = cfg.get("per_val", 10), cfg.get("per_write", 50)
        self.epoch, self.iter = start_epoch, start_iter
        self.log_lag = max(0, min(int(log_lag), self.RING - 2))
        self.dev = next(model.parameters()).device
        self.process_group = process_group
        self.step_size_up = max(1, cfg.max_epoch * cfg.num_mini_batch_per_epoch // 6)
        self.optimizer = None   # FlatAdam, built after the bucket-discovery step (needs the set of parameters that receive gradients)
        self._graph = None
        self._pending = []      # (slot, event, info) of steps whose losses have not been read yet
        self._copy_stream = torch.cuda.Stream(self.dev)

    # ---- schedules (host side, closed form) -------------------------------------------------------------------------------------
    def current_lr(self, it=None):
        return cyclic_lr(self.iter if it is None else it, 1e-5, 1e-3, self.step_size_up)

    def current_bn_momentum(self, it=None):
        bn = self.cfg.bn
        return bn_momentum_at(self.iter if it is None else it, bn.bn_momentum, bn.bn_decay, bn.decay_step, bn.bnm_clip)

    def _apply_schedules(self):
        lr, mom = self.current_lr(), self.current_bn_momentum()
        for g in self.optimizer.param_groups:
            g["lr"] = lr  # FlatAdam.sync_lr (before_replay) mirrors it into the device scalar the captured Adam kernel reads
        if mom != getattr(self, "_mom", None):
            self._mom = mom
            for m in self._bns:
                m.momentum = mom  # GraphedTrainStep refreshes the device momentum table before the replay
        return lr

    # ---- one-time construction of the captured step ------------------------------------------------------------------------------
    def _loss_fn(self, b1, b2):
        def fn(ep):
            syn = {k: v[:b1] for k, v in ep.items()}
            real = {k: v[b1:] for k, v in ep.items()}
            ls, lr = self.loss["syn"](syn), self.loss["real"](real)
            la = (ls * b1 + lr * b2) / (b1 + b2)  # utils/solver.py:182
            self._loss3 = torch.stack([la, ls, lr]).detach()
            return la

        return fn

    def _build(self, syn, real):
        from .graph import GraphedTrainStep
        from .parallel import DataParallelStep, FlatAdam, GradAllReducer, broadcast_module

        b1, b2 = syn["rgb"].shape[0], real["rgb"].shape[0]
        keys = tuple(dict.fromkeys(MODEL_IN + LABELS))
        self._keys, self._b = keys, (b1, b2)
        shape = lambda k: (b1 + b2,) + tuple(syn[k].shape[1:])
        self._host = {k: torch.empty(shape(k), dtype=syn[k].dtype).pin_memory() for k in keys}
        self._stage = {k: torch.empty(shape(k), dtype=syn[k].dtype, device=self.dev) for k in keys}
        merge_into(self._host, syn, real, keys)
        example = {k: v.to(self.dev) for k, v in self._host.items()}
        self._bns = [m for m in self.model.modules() if isinstance(m, torch.nn.modules.batchnorm._BatchNorm)]
        self._mom = None
        mom = self.current_bn_momentum()
        for m in self._bns:
            m.momentum = mom
        self._mom = mom
        broadcast_module(self.model, process_group=self.process_group)
        self.reducer = GradAllReducer(self.model, process_group=self.process_group)
        loss_fn = self._loss_fn(b1, b2)
        # discovery step: which parameters receive gradients (freeze_world_enhancer: utils/solver.py:40-44 filters on requires_grad).
        # It must not leave a trace in the training state, so the BatchNorm buffers are restored afterwards.
        saved = {n: b.detach().clone() for n, b in self.model.named_buffers()}
        self.reducer.zero_grad()
        ep = self.model({k: example[k] for k in MODEL_IN})
        ep.update({k: example[k] for k in LABELS})
        loss_fn(ep).backward()
        self.reducer.finish()
        with torch.no_grad():
            for n, b in self.model.named_buffers():
                b.copy_(saved[n])
        opt_cfg = self.cfg.optimizer
        self.optimizer = FlatAdam(self.reducer, lr=self.current_lr(), weight_decay=opt_cfg.get("weight_decay", 0.0))
        if getattr(self, "_optimizer_state", None) is not None:
            self.load_optimizer_state(self._optimizer_state)
        nccl_in_graph = os.environ.get("ISTNET_GRAPH_NCCL", "0") == "1"
        self._dp = DataParallelStep(self.reducer, self.optimizer, nccl_in_graph=nccl_in_graph)
        saved = {n: b.detach().clone() for n, b in self.model.named_buffers()}
        flat0 = [(f["p"].clone(), f["m"].clone(), f["v"].clone()) for f in self.optimizer.flat]
        step0 = self.optimizer.step_dev.clone()
        self._graph = GraphedTrainStep(self.model, loss_fn, example, MODEL_IN, LABELS, before_forward=self._dp.before, after_backward=self._dp.after,
                                       before_replay=self._dp.before_replay, after_replay=self._dp.after_replay)
        # warm-up and capture ran optimizer steps and BatchNorm updates on the example batch: rewind, the first replay is iteration `iter`
        with torch.no_grad():
            for n, b in self.model.named_buffers():
                b.copy_(saved[n])
            for f, (p, m, v) in zip(self.optimizer.flat, flat0):
                f["p"].copy_(p), f["m"].copy_(m), f["v"].copy_(v)
            self.optimizer.step_dev.copy_(step0)
        from . import nhwc

        nhwc.invalidate_weights(self.dev)
        self._loss_pinned = torch.zeros(self.RING, 3, dtype=torch.float32).pin_memory()
        self._loss_events = [torch.cuda.Event() for _ in range(self.RING)]
        self._staged = torch.cuda.Event()
        self._consumed = torch.cuda.Event()
        self._consumed.record()

    # ---- checkpoint helpers (utils/solver.py:64-68 stores optimizer.state_dict(); FlatAdam's state lives in the flat buffers) -----
    def optimizer_state(self):
        o = self.optimizer
        return {"step": int(o.step_dev.item()), "flat": [{k: f[k].detach().cpu() for k in ("m", "v")} for f in o.flat]}

    def load_optimizer_state(self, state):
        if self.optimizer is None:
            self._optimizer_state = state
            return
        o = self.optimizer
        o.step_dev.fill_(int(state["step"]))
        for f, s in zip(o.flat, state["flat"]):
            f["m"].copy_(s["m"]), f["v"].copy_(s["v"])

    # ---- the loop ---------------------------------------------------------------------------------------------------------------------
    def _stage_batch(self, syn, real):
        """Host: syn | real into the pinned set; copy stream: pinned -> device staging (waits until the previous staging was consumed)."""
        b1, b2 = self._b
        if syn["rgb"].shape[0] != b1 or real["rgb"].shape[0] != b2:
            raise ValueError(f"istnet_b200.solver: the captured step is static ({b1} syn + {b2} real instances); got "
                             f"{syn['rgb'].shape[0]} + {real['rgb'].shape[0]} (use drop_last=True as config/ist_net_default.yaml does)")
        self._consumed.synchronize()  # the previous batch has left `_stage` (and, before that, `_host`): both may be rewritten
        merge_into(self._host, syn, real, self._keys)
        with torch.cuda.stream(self._copy_stream):
            for k in self._keys:
                self._stage[k].copy_(self._host[k], non_blocking=True)
            self._staged.record(self._copy_stream)

    def _drain(self, keep):
        """Moves finished steps' losses from the pinned ring into the log buffer, leaving at most `keep` steps in flight."""
        while len(self._pending) > keep:
            slot, info = self._pending.pop(0)
            self._loss_events[slot].synchronize()
            la, ls, lr = (float(x) for x in self._loss_pinned[slot])
            d = {"loss_all": la, "loss_syn": ls, "loss_real": lr}
            d.update(info)
            self.log_buffer.update(d)

    def step(self, syn_data, real_data, mode="train"):
        """One iteration: returns the device tensor [loss_all, loss_syn, loss_real] of THIS step (asynchronous)."""
        if self._graph is None:
            self._build(syn_data, real_data)
        lr = self._apply_schedules()
        self._stage_batch(syn_data, real_data)
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_event(self._staged)
        self._graph.load(self._stage)
        self._consumed.record(cur)
        self._graph()
        slot = self._slot = self.iter % self.RING
        self._loss_pinned[slot].copy_(self._loss3, non_blocking=True)
        self._loss_events[slot].record(cur)
        return self._loss3, {"lr": lr}

    def train(self):
        self.model.train()
        end = time.time()
        for name in ("syn", "real"):
            ds = getattr(self.dataloaders[name], "dataset", None)
            if hasattr(ds, "reset"):
                ds.reset()  # utils/solver.py:80-81
        n_iter = len(self.dataloaders["syn"]) if hasattr(self.dataloaders["syn"], "__len__") else -1
        i = 0
        for syn_data, real_data in zip(self.dataloaders["syn"], self.dataloaders["real"]):
            data_time = time.time() - end
            _, info = self.step(syn_data, real_data)
            info.update({"T_data": data_time, "T_step": time.time() - end - data_time})
            self._pending.append((self._slot, info))
            self._drain(self.log_lag)
            if i % self.per_write == 0 and self.log_buffer.val_history:
                self.log_buffer.average(self.per_write)
                prefix = "[{}/{}][{}/{}][{}] Train - ".format(self.epoch, self.cfg.max_epoch, i, n_iter, self.iter)
                if self.logger is not None:
                    self.logger.info(self.get_logger_info(prefix, dict_info=self.log_buffer._output))
                self.write_summary(self.log_buffer._output, "train")
            end = time.time()
            self.iter += 1
            i += 1
        self._drain(0)
        out = self.log_buffer.avg
        self.log_buffer.clear()
        return out

    def solve(self, save_checkpoint=None):
        """utils/solver.py:51-72.  `save_checkpoint(model, filename, optimizer, meta)`: gorilla.solver.save_checkpoint or compatible."""
        while self.epoch <= self.cfg.max_epoch:
            if self.logger is not None:
                self.logger.info("\nEpoch {} :".format(self.epoch))
            end = time.time()
            info = self.train()
            d = {"train_time(min)": (time.time() - end) / 60.0}
            d.update({"train_" + k: v for k, v in info.items() if "loss" in k})
            if self.epoch % 5 == 0 and save_checkpoint is not None:
                path = os.path.join(self.cfg.log_dir, "epoch_" + str(self.epoch) + ".pth")
                save_checkpoint(model=self.model, filename=path, optimizer=None, meta={"iter": self.iter, "epoch": self.epoch,
                                                                                       "flat_adam": self.optimizer_state()})
            if self.logger is not None:
                self.logger.warning(self.get_logger_info("Epoch {} - ".format(self.epoch), dict_info=d))
            self.epoch += 1

    def get_logger_info(self, prefix, dict_info):
        info = prefix
        for key, value in dict_info.items():
            info += ("{}: {:.3f}\t" if "T_" in key else "{}: {:.5f}\t").format(key, value)
        return info

    def write_summary(self, dict_info, mode):
        if self.tb_writer is None:
            return
        assert mode in ("train", "eval")
        if hasattr(self.tb_writer, "update_scalar"):  # the reference's tools_writer (utils/solver.py:243-275)
            self.tb_writer.update_scalar(list_name=list(dict_info.keys()), list_value=list(dict_info.values()),
                                         index_counter=0 if mode == "train" else 1, prefix=mode + "_")
        else:
            for k, v in dict_info.items():
                self.tb_writer.add_scalar(mode + "_" + k, v, self.iter)
