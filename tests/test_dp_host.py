"""Data-parallel host logic (dp.py) on CPU: world size 2 over gloo, spawned with torch.multiprocessing.

The rank-local "step" is the CPU oracle (oracle/gscan_oracle.py, float64 autograd) on the rank's shard; what
is under test is everything dp.py adds around it: the shard bounds, the SUM-form loss, the counts appended to
the flat gradient buffer, the ONE all-reduce, the division by the global token count - the result must equal
the gradient of the reference's global-batch loss (train.py:102-107 with the normalisations of
model.py:100,159,163) computed by a single process on the concatenated batch.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multimodal_seq2seq_gscan_b200 import dp
from oracle import gscan_oracle as O

W_AUX = 0.3


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _cfg(aux):
    cfg = dict(O.CONFIGS["tiny"])
    cfg["auxiliary_task"] = aux
    return cfg


def _global_batch(cfg, B):
    return O.synthetic_batch(cfg, batch_size=B, seed=77, max_tgt_len=9)


def _as_iterator_tuple(batch):
    """The 9-tuple GroundedScanDataset.get_data_iterator yields (gSCAN_dataset.py:229-231)."""
    return (torch.tensor(batch["commands"]), np.asarray(batch["cmd_lengths"], dtype=np.float64), None,
            torch.tensor(batch["situations"]), None, torch.tensor(batch["targets"]),
            np.asarray(batch["tgt_lengths"], dtype=np.float64), None, torch.tensor(batch["target_positions"]))


def _flat(grads, names):
    return torch.cat([grads[k].reshape(-1) for k in names])


def _single_process_gradient(cfg, params, batch):
    p = {k: v.detach().double().requires_grad_(True) for k, v in params.items()}
    logp, aux = O.model_forward(p, torch.tensor(batch["commands"]), batch["cmd_lengths"],
                                torch.tensor(batch["situations"]).double(), torch.tensor(batch["targets"]),
                                cfg["conditional_attention"], cfg["auxiliary_task"])
    loss = O.nll_loss(logp, torch.tensor(batch["targets"]))
    if cfg["auxiliary_task"]:
        loss = loss + W_AUX * O.aux_nll_loss(aux, torch.tensor(batch["target_positions"]))
    loss.backward()
    return loss.detach(), {k: v.grad.detach() for k, v in p.items()}


def _worker(rank, world, port, aux, B, know_global, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        cfg = _cfg(aux)
        params = O.synthetic_params(cfg, 5, scale=2.0)
        names = [n for n, _ in O.param_shapes(cfg)]
        sharded = dp.shard_batch(_as_iterator_tuple(_global_batch(cfg, B)), rank, world, aux)
        if sharded is None:
            # every rank must take this branch together (B < world): nothing to exchange, nothing hangs
            torch.save({"skipped": True}, os.path.join(out_dir, f"r{rank}.pt"))
            return
        (cmds, cmd_len, _, sits, _, tgts, tgt_len, _, pos), global_counts = sharded
        p = {k: v.detach().double().requires_grad_(True) for k, v in params.items()}
        counts_work = None
        gc = global_counts if know_global else None
        if aux and gc is None:
            gc, counts_work = dp.start_count_allreduce(tgts, 0)
        logp, aux_logp = O.model_forward(p, cmds, [int(x) for x in cmd_len], sits.double(), tgts,
                                         cfg["conditional_attention"], aux)
        nll = O.nll_loss(logp, tgts)
        n_tok = (tgts[:, 1:] != 0).sum().double()
        aux_mean = O.aux_nll_loss(aux_logp, pos) if aux else None
        if counts_work is not None:
            counts_work.wait()
        loss = dp.sum_loss(nll, n_tok, aux_mean, tgts.shape[0], W_AUX, gc)
        loss.backward()
        n = sum(int(np.prod(s)) for _, s in O.param_shapes(cfg))
        flat = torch.zeros(n + dp.COUNT_SLOTS, dtype=torch.float64)
        flat[:n] = _flat({k: v.grad for k, v in p.items()}, names)
        dp.pack_counts(flat, n, dp.local_counts(tgts, 0).double())
        dp.allreduce_flat_gradient(flat)            # the single gradient collective
        grad = flat[:n] / flat[n]                   # what gscan_adam_step_dev does on the device
        torch.save({"grad": grad, "counts": flat[n:n + 2].clone(), "loss_share": loss.detach() / flat[n],
                    "shard": (int(tgts.shape[0]), int(tgts.shape[1]), int(cmds.shape[1])),
                    "global_counts": global_counts}, os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def _run(world, aux, B, know_global, tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, aux, B, know_global, str(tmp_path)), nprocs=world, join=True)
    return [torch.load(os.path.join(tmp_path, f"r{r}.pt"), weights_only=False) for r in range(world)]


def test_shard_bounds_cover_everything_once():
    for n in (1, 2, 3, 7, 8, 200, 201):
        for world in (1, 2, 3, 8):
            spans = [dp.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("aux,know_global", [(False, True), (False, False), (True, True), (True, False)])
@pytest.mark.parametrize("B", [7, 8])
def test_allreduced_gradient_equals_single_process(aux, know_global, B, tmp_path):
    """Ragged (7 = 4 + 3) and even shards, auxiliary task off / on, global counts known to the caller or exchanged."""
    cfg = _cfg(aux)
    params = O.synthetic_params(cfg, 5, scale=2.0)
    names = [n for n, _ in O.param_shapes(cfg)]
    batch = _global_batch(cfg, B)
    loss_ref, grads_ref = _single_process_gradient(cfg, params, batch)
    ref = _flat(grads_ref, names)
    outs = _run(2, aux, B, know_global, tmp_path)
    n_tok = float(sum(l - 1 for l in batch["tgt_lengths"]))
    for r, o in enumerate(outs):
        assert o["counts"].tolist() == [n_tok, float(B)]
        assert o["global_counts"] == (n_tok, float(B))
        rel = ((o["grad"] - ref).norm() / ref.norm()).item()
        assert rel < 1e-12, (r, rel)
    # identical replicas: both ranks hold bit-identical gradients after the collective
    assert torch.equal(outs[0]["grad"], outs[1]["grad"])
    # the loss shares add up to the global-batch loss
    total = sum(o["loss_share"] for o in outs)
    assert abs(total.item() - loss_ref.item()) < 1e-12 * max(1.0, abs(loss_ref.item()))
    # shard shapes: contiguous halves; targets keep the global padded length only when the aux task needs it
    sizes = [o["shard"][0] for o in outs]
    assert sizes == [B - B // 2, B // 2]
    if aux:
        assert all(o["shard"][1] == batch["targets"].shape[1] for o in outs)


def test_batch_smaller_than_world_is_skipped_by_every_rank(tmp_path):
    """n_train % batch_size < world must not leave some ranks in a collective (train.py)."""
    outs = _run(2, False, 1, True, tmp_path)
    assert all(o.get("skipped") for o in outs)
