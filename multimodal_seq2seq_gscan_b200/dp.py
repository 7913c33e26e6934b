"""Batch-sharded data parallelism for the training step (new; the reference is single-device).

Examples are independent in forward and backward; the only coupling between examples is the loss
normalisation: ``NLLLoss(ignore_index)`` divides by the GLOBAL number of non-pad target tokens
(reference model.py:100,159) and the auxiliary loss by the global batch size (model.py:59,163).

One collective per step.  Every rank backpropagates the SUM form of its shard's loss

    L_r = nll_sum_r + c * aux_sum_r                       (nll_sum_r = nll_mean_r * n_tok_r, ...)

into one flat fp32 buffer laid out in ``model.parameters()`` order with ``[n_tok_r, B_r]`` appended
(``COUNT_SLOTS`` floats after the last gradient).  ONE all-reduce (SUM) of that buffer - NCCL over
NVLink on the GPUs, gloo in the CPU tests - yields ``sum_r dL_r`` and ``[N_tok, B_all]``; the fused
Adam kernel divides by ``N_tok`` as it reads the gradient (``gscan_adam_step_dev``), so

    g = (sum_r d nll_sum_r) / N_tok + w * (sum_r d aux_sum_r) / B_all

which is the gradient of the reference's global-batch loss ``nll_mean + w * aux_mean``
(train.py:102-107) when ``c = w * N_tok / B_all``.

``c`` needs the global token count BEFORE the backward pass.  It is known without communication
whenever the caller holds the global batch (``train.py``: every rank builds the same shuffled
batches and slices its shard; ``bench.py``: synthetic shards of every rank are generated from the
rank number) - pass it as ``global_counts``.  Without the auxiliary task ``c`` is not needed at all.
Only for "auxiliary task on AND global counts unknown" a second, tiny all-reduce of the two counts
is started before the forward pass (which hides it) - ``start_count_allreduce``.

Nothing here touches CUDA directly, which is what lets the host logic be tested with gloo
(tests/test_dp_host.py) against the CPU oracle.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

# floats appended to the flat gradient buffer: [n_tok, n_examples, 0, 0] (16-byte granularity)
COUNT_SLOTS = 4


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n examples for `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(batch: Sequence, rank: int, world: int, auxiliary_task: bool):
    """This rank's contiguous shard of one global batch of ``GroundedScanDataset.get_data_iterator``
    (the 9-tuple of gSCAN_dataset.py:229-231), or ``None`` when the global batch has fewer examples
    than there are ranks - EVERY rank then skips it (same decision everywhere: no rank is left
    waiting in a collective; the last, ragged batch of an epoch is where this happens).

    Commands are cut to the shard's own maximal length (masked attention: no effect on the result);
    targets too unless the auxiliary task is on, whose scores sum the visual attention over EVERY
    padded step of the global batch (SURVEY.md trap A.4-2).  Also returns the global counts
    ``(N_tok, B_all)`` as Python numbers, computed from the global target lengths every rank holds (pad index 0,
    as in every gSCAN vocabulary: the scored targets of an example are its tokens after SOS)."""
    (input_batch, input_lengths, derivation, situation_batch, situation_repr, target_batch, target_lengths,
     agent_positions, target_positions) = batch
    n = int(input_batch.shape[0])
    if n < world:
        return None
    lo, hi = shard_bounds(n, rank, world)
    tl_all = np.asarray(target_lengths).reshape(-1)
    # scored targets per example: everything after SOS up to and including EOS (model.py:108-115,159)
    global_counts = (float(np.sum(tl_all - 1)), float(n))
    il, tl = input_lengths[lo:hi], target_lengths[lo:hi]
    ib = input_batch[lo:hi, :int(np.max(np.asarray(il)))]
    tb = target_batch[lo:hi] if auxiliary_task else target_batch[lo:hi, :int(np.max(np.asarray(tl)))]
    sub = (ib, il, None if derivation is None else derivation[lo:hi], situation_batch[lo:hi],
           None if situation_repr is None else situation_repr[lo:hi], tb, tl,
           None if agent_positions is None else agent_positions[lo:hi], target_positions[lo:hi])
    return sub, global_counts


def batch_const(targets: torch.Tensor) -> torch.Tensor:
    """The number of examples of ``targets`` as a cached 1-element fp32 tensor on its device (no host-to-device copy
    per step: such an assignment stalls the enqueueing thread until the stream drains)."""
    key = (targets.device, int(targets.shape[0]), 1)
    t = _BATCH_CONST.get(key)
    if t is None:
        t = _BATCH_CONST[key] = torch.full((1,), float(targets.shape[0]), dtype=torch.float32, device=targets.device)
    return t


def local_counts(targets: torch.Tensor, pad_idx: int) -> torch.Tensor:
    """[non-pad target tokens after the SOS column, examples] of this rank's shard, on the device of
    `targets` - exactly what NLLLoss(ignore_index) / the auxiliary NLLLoss divide by (model.py:100,59)."""
    # (no `counts[i] = python_float`: that assignment goes through a host-to-device copy that stalls the
    #  enqueueing thread until the stream drains - measured +0.3 ms per step on B200)
    key = (targets.device, int(targets.shape[0]))
    if key not in _BATCH_CONST:
        _BATCH_CONST[key] = torch.full((), float(targets.shape[0]), dtype=torch.float32, device=targets.device)
    n_tok = (targets[:, 1:] != pad_idx).sum(dtype=torch.float32)
    if pad_idx != 0:      # the literal 0 appended behind every shifted sequence is scored then (model.py:108-115)
        n_tok = n_tok + _BATCH_CONST[key]
    return torch.stack((n_tok, _BATCH_CONST[key]))


_BATCH_CONST: dict = {}


def start_count_allreduce(targets: torch.Tensor, pad_idx: int, group=None):
    """Only for "auxiliary task on and global counts unknown": the all-reduce of the two counts, started
    BEFORE the forward pass (they depend on the targets only) and waited for when the loss is formed.
    Returns (tensor, work)."""
    counts = local_counts(targets, pad_idx)
    work = dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group, async_op=True)
    return counts, work


def sum_loss(nll_mean: torch.Tensor, n_tokens: torch.Tensor, aux_mean: Optional[torch.Tensor], batch_size: int,
             weight_target_loss: float, global_counts) -> torch.Tensor:
    """The SUM-form loss of this rank's shard (module docstring): its gradients, summed over ranks and
    divided by the global token count, are the gradients of the reference's global-batch loss.
    ``global_counts`` = (N_tok, B_all) as numbers or a 2-element tensor; may be None without the
    auxiliary task."""
    loss = nll_mean * n_tokens.detach()
    if aux_mean is not None:
        if global_counts is None:
            raise ValueError("the auxiliary loss needs the global counts before the backward pass")
        n_tok_all, b_all = global_counts[0], global_counts[1]
        if isinstance(n_tok_all, torch.Tensor):      # exchanged counts: fp32 integers, exact; their ratio in fp64
            ratio = (n_tok_all.double() / b_all.double()).to(aux_mean.dtype)
        else:
            ratio = n_tok_all / b_all
        loss = loss + aux_mean * (weight_target_loss * batch_size * ratio)
    return loss


def global_loss(nll_mean: torch.Tensor, aux_mean: Optional[torch.Tensor], weight_target_loss: float) -> torch.Tensor:
    """Single process: exactly train.py:102-107."""
    loss = nll_mean
    if aux_mean is not None:
        loss = loss + weight_target_loss * aux_mean
    return loss


def pack_counts(flat_grad: torch.Tensor, n_grad: int, counts: torch.Tensor) -> None:
    """Write this rank's [n_tok, n_examples] behind the gradients of the flat buffer."""
    flat_grad[n_grad:n_grad + 2].copy_(counts)


def allreduce_flat_gradient(flat_grad: torch.Tensor, group=None) -> None:
    """The single collective of the step: SUM over ranks of [gradients | n_tok | n_examples], in place."""
    dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
