"""Fused training step: the body of the reference's training loop (train.py:96-114) on the B200
kernels, with the parameters, gradients and Adam state each held in ONE flat fp32 buffer so that the
optimizer is a single kernel launch and data parallelism is a single all-reduce.

The reference's own ``train.py`` also runs unchanged against ``Model`` (torch.optim.Adam over
``model.parameters()``); this class is the fast path and the data-parallel path.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch
import torch.distributed as dist

from . import dp, ops


class FusedTrainer:
    def __init__(self, model, learning_rate: float = 1e-3, adam_beta_1: float = 0.9, adam_beta_2: float = 0.999,
                 lr_decay: float = 0.9, lr_decay_steps: int = 20000, weight_target_loss: float = 0.3,
                 eps: float = 1e-8, distributed: bool = False, process_group=None):
        self.model = model
        self.lr, self.betas, self.eps = learning_rate, (adam_beta_1, adam_beta_2), eps
        self.lr_decay, self.lr_decay_steps = lr_decay, lr_decay_steps
        self.weight_target_loss = weight_target_loss
        self.distributed = distributed
        self.group = process_group
        self.step_count = 0
        self.params = list(model.parameters())
        plist = model._param_list()
        self._shapes = [None if p is None else tuple(p.shape) for p in plist]
        self._present = [p for p in plist if p is not None]
        assert [id(p) for p in self._present] == [id(p) for p in self.params]
        sizes, offsets = ops.flat_layout(self._shapes)
        self._sizes, self._offsets = sizes, offsets
        dev = self.params[0].device
        n = int(offsets[-1])
        self.flat_param = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        # re-seat every parameter as a view into the flat buffer (identity of the nn.Parameter kept)
        with torch.no_grad():
            for p, view in zip(self._present, self._views(self.flat_param)):
                view.copy_(p.data)
                p.data = view
        if distributed:
            # identical replicas: rank 0's parameters win
            dist.broadcast(self.flat_param, src=0, group=process_group)

    def _views(self, flat):
        return [flat[int(o):int(o) + n].view(s) for s, o, n in zip(self._shapes, self._offsets[:-1], self._sizes)
                if s is not None]

    def current_lr(self) -> float:
        """LambdaLR(lr_decay ** (t / lr_decay_steps)) with t = optimizer steps taken so far (train.py:69-70)."""
        return self.lr * self.lr_decay ** (self.step_count / self.lr_decay_steps)

    def _flat_gradient(self, grads):
        base = grads[0]._base
        ok = base is not None and base.numel() == self.flat_param.numel()
        if ok:
            for g, o in zip(grads, [o for s, o in zip(self._shapes, self._offsets[:-1]) if s is not None]):
                if g._base is not base or g.storage_offset() != int(o):
                    ok = False
                    break
        if ok:
            return base
        flat = torch.zeros_like(self.flat_param)
        for g, view in zip(grads, self._views(flat)):
            view.copy_(g)
        return flat

    def train_step(self, commands, commands_lengths, situations, targets, target_lengths, target_positions=None):
        """One iteration: forward, loss, backward, (gradient all-reduce), Adam, LR schedule.
        Returns the (global-batch) loss share of this rank as a 0-dim device tensor - no host sync."""
        model = self.model
        model.train()
        counts = work = None
        if self.distributed:   # depends on the targets only: in flight during the forward pass
            counts, work = dp.start_count_allreduce(targets, model.target_pad_idx, self.group)
        logp, aux = model(commands_input=commands, commands_lengths=commands_lengths, situations_input=situations,
                          target_batch=targets, target_lengths=target_lengths)
        nll, n_tok = ops.NLLLoss.apply(logp, targets, model.target_pad_idx, 1)
        aux_mean = None
        if model.auxiliary_task and target_positions is not None and self.weight_target_loss != 0:
            aux_mean = model.get_auxiliary_loss(aux, target_positions)
        if work is not None:
            work.wait()
        loss = dp.global_loss(nll, n_tok, aux_mean, targets.shape[0], self.weight_target_loss, counts)
        grads = torch.autograd.grad(loss, self._present)
        flat_grad = self._flat_gradient(grads)
        if self.distributed:
            dp.allreduce_flat_gradient(flat_grad, self.group)
        lr = self.current_lr()
        self.step_count += 1
        ops.adam_step(self.flat_param, flat_grad, self.exp_avg, self.exp_avg_sq, lr, self.betas[0], self.betas[1],
                      self.eps, self.step_count)
        model.update_state(is_best=False)
        self.last_logp, self.last_aux = logp.detach(), aux
        return loss.detach()

    # ---- torch.optim.Adam-compatible state, so checkpoints interoperate with the reference -------
    def state_dict(self) -> dict:
        state = {}
        for i, (m, v) in enumerate(zip(self._views(self.exp_avg), self._views(self.exp_avg_sq))):
            state[i] = {"step": torch.tensor(float(self.step_count)), "exp_avg": m.clone(), "exp_avg_sq": v.clone()}
        group = {"lr": self.current_lr(), "betas": self.betas, "eps": self.eps, "weight_decay": 0, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "decoupled_weight_decay": False, "initial_lr": self.lr, "params": list(range(len(self._present)))}
        return {"state": state if self.step_count else {}, "param_groups": [group]}

    def load_state_dict(self, sd: dict) -> None:
        if sd["state"]:
            for i, (m, v) in enumerate(zip(self._views(self.exp_avg), self._views(self.exp_avg_sq))):
                m.copy_(sd["state"][i]["exp_avg"])
                v.copy_(sd["state"][i]["exp_avg_sq"])
            self.step_count = int(sd["state"][0]["step"])
