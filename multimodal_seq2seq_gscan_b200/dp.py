"""Batch-sharded data parallelism for the training step (new; the reference is single-device).

Examples are independent in forward and backward; the only coupling between examples is the loss
normalisation: ``NLLLoss(ignore_index)`` divides by the GLOBAL number of non-pad target tokens
(reference model.py:100,159) and the auxiliary loss by the global batch size (model.py:59,163).
So each rank runs the ordinary step on its shard with its loss re-weighted to

    loss_r = nll_mean_r * n_tok_r / N_tok + w * aux_mean_r * B_r / B_all

after which the SUM over ranks of the gradients equals the gradient of the reference's
global-batch loss.  Per step there is one tiny all-reduce of the two counts (started before the
forward pass, which hides it) and ONE all-reduce of the flat fp32 gradient buffer (440,275 floats = 1.76 MB for
the compositional config) - NCCL over NVLink on the GPUs, gloo in the CPU tests.
Nothing here touches CUDA directly, which is what lets the host logic be tested with gloo.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n examples for `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def local_counts(targets: torch.Tensor, pad_idx: int) -> torch.Tensor:
    """[non-pad target tokens after the SOS column, examples] of this rank's shard, on the device of
    `targets` - exactly what NLLLoss(ignore_index) / the auxiliary NLLLoss divide by (model.py:100,59)."""
    # (no `counts[i] = python_float`: that assignment goes through a host-to-device copy that stalls the
    #  enqueueing thread until the stream drains - measured +0.3 ms per step on B200)
    key = (targets.device, int(targets.shape[0]))
    if key not in _BATCH_CONST:
        _BATCH_CONST[key] = torch.full((), float(targets.shape[0]), dtype=torch.float32, device=targets.device)
    return torch.stack(((targets[:, 1:] != pad_idx).sum(dtype=torch.float32), _BATCH_CONST[key]))


_BATCH_CONST: dict = {}


def start_count_allreduce(targets: torch.Tensor, pad_idx: int, group=None):
    """Launch the all-reduce of the two counts; they depend on the targets only, so the collective is
    started BEFORE the forward pass and waited for when the loss is formed.  Returns (tensor, work)."""
    counts = local_counts(targets, pad_idx)
    work = dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group, async_op=True)
    return counts, work


def global_loss(nll_mean: torch.Tensor, n_tokens: torch.Tensor, aux_mean: Optional[torch.Tensor], batch_size: int,
                weight_target_loss: float, global_counts: Optional[torch.Tensor]) -> torch.Tensor:
    """This rank's share of the global-batch loss (see module docstring).  With
    ``global_counts=None`` (single process) it is exactly train.py:102-107."""
    if global_counts is None:
        loss = nll_mean
        if aux_mean is not None:
            loss = loss + weight_target_loss * aux_mean
        return loss
    loss = nll_mean * (n_tokens.detach() / global_counts[0])
    if aux_mean is not None:
        loss = loss + aux_mean * (weight_target_loss * batch_size / global_counts[1])
    return loss


def allreduce_flat_gradient(flat_grad: torch.Tensor, group=None) -> None:
    """The single gradient collective of the step: SUM over ranks, in place."""
    dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
