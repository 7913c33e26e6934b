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
        self.collective_events = None    # set to a list to have (before, after) CUDA events of the gradient all-reduce appended
        self.params = list(model.parameters())
        plist = model._param_list()
        self._shapes = [None if p is None else tuple(p.shape) for p in plist]
        self._present = [p for p in plist if p is not None]
        assert [id(p) for p in self._present] == [id(p) for p in self.params]
        sizes, offsets = ops.flat_layout(self._shapes)
        self._sizes, self._offsets = sizes, offsets
        dev = self.params[0].device
        n = int(offsets[-1])
        self._n = n
        self.flat_param = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        # persistent flat gradient buffer: [gradients in parameters() order | n_tok, n_examples, 0, 0]; the backward
        # pass writes into it (ops.set_flat_grad_target), the data-parallel all-reduce sums all of it at once
        self.flat_grad = torch.zeros(n + dp.COUNT_SLOTS, dtype=torch.float32, device=dev)
        self._side_stream = torch.cuda.Stream(device=dev)   # zeroing, early loss gradient, loss value (train_step)
        self._d_logp = None
        self._count_out = None
        self._d_aux = None
        self._aux_count_out = None
        self._aux_scale = None
        self._one = torch.ones((), dtype=torch.float32, device=dev)   # d loss / d loss, without a fill kernel per step
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

    @staticmethod
    def _aux_cells(model, situations) -> int:
        """Number of grid cells the auxiliary head scores (model.py:166-170: one log-probability per cell)."""
        return int(situations.shape[1] * situations.shape[2])

    def current_lr(self) -> float:
        """LambdaLR(lr_decay ** (t / lr_decay_steps)) with t = optimizer steps taken so far (train.py:69-70)."""
        return self.lr * self.lr_decay ** (self.step_count / self.lr_decay_steps)

    def _flat_gradient(self, grads):
        """The flat buffer the per-parameter gradients are views of (what ``ModelForward.backward`` produced), or a
        packed copy when they came from somewhere else."""
        base = grads[0]._base
        ok = base is not None and base.numel() == self._n + dp.COUNT_SLOTS
        if ok:
            for g, o in zip(grads, [o for s, o in zip(self._shapes, self._offsets[:-1]) if s is not None]):
                if g._base is not base or g.storage_offset() != int(o):
                    ok = False
                    break
        if ok:
            return base
        flat = torch.zeros(self._n + dp.COUNT_SLOTS, dtype=torch.float32, device=self.flat_param.device)
        for g, view in zip(grads, self._views(flat)):
            view.copy_(g)
        return flat

    def train_step(self, commands, commands_lengths, situations, targets, target_lengths, target_positions=None,
                   global_counts=None):
        """One iteration: forward, loss, backward, (ONE all-reduce of gradients + counts), Adam, LR schedule.
        Returns this rank's share of the global-batch loss as a 0-dim device tensor (the shares of all ranks add
        up to the reference's loss of the global batch) - no host sync.

        ``global_counts`` = (non-pad target tokens, examples) of the GLOBAL batch as numbers, when the caller knows
        them (train.py and bench.py do): needed before the backward pass only by the auxiliary loss - without
        them and with the auxiliary task on, a small count all-reduce runs beside the forward pass (dp.py)."""
        model = self.model
        if not model.training:           # nn.Module.train() walks every submodule: 0.2 ms of host time per step
            model.train()
        n = self._n
        use_aux = bool(model.auxiliary_task and target_positions is not None and self.weight_target_loss != 0)
        counts_work = None
        if self.distributed and use_aux:
            # this rank's [n_tok, n_examples] go behind the gradients NOW: they depend on the targets only
            local = dp.local_counts(targets, model.target_pad_idx)
            self.flat_grad[n:n + 2].copy_(local)
            if global_counts is None:
                global_counts, counts_work = dp.start_count_allreduce(targets, model.target_pad_idx, self.group)
        # Off the critical path, on a second stream beside the forward pass: the zeroing of the gradient buffer
        # (ModelForward.backward is told so below) and - without the auxiliary task - d(loss)/d(logp), which depends on the
        # targets only (-1/N_tok, resp. -1 in SUM form, at the scored positions).  Between the two sweeps of the step
        # there is then nothing but the output head; the loss VALUE (reporting only) is computed on that stream too,
        # beside the backward pass.
        main = torch.cuda.current_stream(self.flat_param.device)
        side = self._side_stream
        # (with the auxiliary task the SUM form needs the ratio N_tok / B of the GLOBAL batch: known in advance when the
        #  caller passed the counts as numbers, else the plain autograd path runs)
        early_grad = (not use_aux) or (not self.distributed) or \
            (global_counts is not None and not isinstance(global_counts[0], torch.Tensor))
        side.wait_stream(main)           # the previous step's Adam has read the gradient buffer; the targets are there
        # (nothing is ALLOCATED under the second stream: blocks that cross streams come back to the caching allocator
        #  late, and the cudaMalloc calls that then fill the gap showed up as 40-80 ms stalls every few dozen steps)
        if early_grad:
            V = int(model._static_cfg["V"])
            if self._d_logp is None or self._d_logp.shape != (targets.shape[0], targets.shape[1], V):
                self._d_logp = torch.empty(targets.shape[0], targets.shape[1], V, dtype=torch.float32, device=targets.device)
                self._count_out = torch.empty(68, dtype=torch.float32, device=targets.device)
            loss_out = torch.empty(68, dtype=torch.float32, device=targets.device)   # fresh: the caller may keep the loss
            if use_aux:
                pos = target_positions.view(-1, 1)
                Bp, Ma = pos.shape[0], self._aux_cells(model, situations)
                if self._d_aux is None or self._d_aux.shape != (Bp, 1, Ma):
                    self._d_aux = torch.empty(Bp, 1, Ma, dtype=torch.float32, device=targets.device)
                    self._aux_count_out = torch.empty(68, dtype=torch.float32, device=targets.device)
                    self._aux_scale = torch.empty(1, dtype=torch.float32, device=targets.device)
                aux_out = torch.empty(68, dtype=torch.float32, device=targets.device)
                # d loss / d (mean auxiliary NLL): its weight, times B_local * N_tok_all / B_all in the SUM form (dp.sum_loss)
                aux_w = float(self.weight_target_loss)
                if self.distributed:
                    aux_w *= targets.shape[0] * (float(global_counts[0]) / float(global_counts[1]))
        with torch.cuda.stream(side):
            self.flat_grad[:n].zero_()
            if early_grad:
                d_logp, _ = ops.nll_grad_from_targets(targets, V, model.target_pad_idx, 1, sum_form=self.distributed,
                                                      d_logp=self._d_logp, out=self._count_out)
                if self.distributed:
                    # this rank's [n_tok, n_examples] behind the gradients (two copies, nothing allocated): the token count
                    # is the one gscan_nll_count has just produced
                    self.flat_grad[n:n + 1].copy_(self._count_out[1:2])
                    self.flat_grad[n + 1:n + 2].copy_(dp.batch_const(targets))
                d_ready = torch.cuda.Event()
                d_ready.record(side)
                if use_aux:      # the auxiliary head's loss gradient, from the target positions alone (model.py:59,162-164)
                    self._aux_scale.fill_(aux_w)
                    d_aux, _ = ops.nll_grad_from_targets(pos, Ma, -100, 0, sum_form=False, d_logp=self._d_aux,
                                                         out=self._aux_count_out, d_loss=self._aux_scale)
        if early_grad:
            # ... and the forward call runs the output-head backward too, most of it beside the decoder sweep
            ops.set_early_dlogp(d_logp, d_ready)
        try:
            logp, aux = model(commands_input=commands, commands_lengths=commands_lengths, situations_input=situations,
                              target_batch=targets, target_lengths=target_lengths)
        finally:
            ops.set_early_dlogp(None)
        main.wait_stream(side)
        if early_grad:
            fwd_done = torch.cuda.Event()
            fwd_done.record(main)
            ops.set_flat_grad_target(self.flat_grad, prezeroed=True)
            try:
                if use_aux:
                    grads = torch.autograd.grad([logp, aux], self._present, grad_outputs=[d_logp, d_aux.view_as(aux)])
                else:
                    grads = torch.autograd.grad([logp], self._present, grad_outputs=[d_logp])
            finally:
                ops.set_flat_grad_target(None)
            aux_mean = None
            with torch.cuda.stream(side), torch.no_grad():
                side.wait_event(fwd_done)
                nll, n_tok = ops.NLLLoss.apply(logp.detach(), targets, model.target_pad_idx, 1, False, loss_out)
                if use_aux:
                    aux_mean = ops.NLLLoss.apply(aux.detach().unsqueeze(1), pos, -100, 0, False, aux_out)[0]
            loss = None      # formed on the caller's stream once it has waited for the second one (below)
        else:
            nll, n_tok = ops.NLLLoss.apply(logp, targets, model.target_pad_idx, 1, False)
            aux_mean = model.get_auxiliary_loss(aux, target_positions)
            if self.distributed:
                if counts_work is not None:
                    counts_work.wait()
                loss = dp.sum_loss(nll, n_tok, aux_mean, targets.shape[0], self.weight_target_loss, global_counts)
            else:
                loss = dp.global_loss(nll, aux_mean, self.weight_target_loss)
            ops.set_flat_grad_target(self.flat_grad, prezeroed=True)
            try:
                grads = torch.autograd.grad(loss, self._present, grad_outputs=self._one)
            finally:
                ops.set_flat_grad_target(None)
        flat_grad = self._flat_gradient(grads)
        denom = None
        if self.distributed:
            if flat_grad is not self.flat_grad:
                flat_grad[n:n + 2].copy_(self.flat_grad[n:n + 2])
            if self.collective_events is not None:
                ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                ev[0].record()
            dp.allreduce_flat_gradient(flat_grad, self.group)     # the single collective of the step
            if self.collective_events is not None:
                ev[1].record()
                self.collective_events.append(ev)
            denom = flat_grad[n:n + 1]                            # global token count, still on the device
        lr = self.current_lr()
        self.step_count += 1
        ops.adam_step(self.flat_param, flat_grad[:n], self.exp_avg, self.exp_avg_sq, lr, self.betas[0], self.betas[1],
                      self.eps, self.step_count, grad_denom=denom)
        main.wait_stream(side)           # the loss value (and nothing else) comes from the second stream
        if loss is None:
            if self.distributed:
                loss = dp.sum_loss(nll, n_tok, aux_mean, targets.shape[0], self.weight_target_loss, global_counts)
            else:
                loss = dp.global_loss(nll, aux_mean, self.weight_target_loss)
        model.update_state(is_best=False)
        self.last_logp, self.last_aux = logp.detach(), aux
        self.last_flat_grad = flat_grad
        return loss.detach() / denom[0] if denom is not None else loss.detach()

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
