"""Data parallelism on real GPUs: 2 NCCL ranks (spawned with torch.multiprocessing) against ONE GPU on the
concatenated batch.  Skipped below 2 visible GPUs (`gpurun --gpus 2 -- python -m pytest tests/test_gpu_dp.py -m gpu`).

What must hold (reference loss normalisation model.py:100,159,163): the all-reduced flat gradient
divided by the global token count, and the parameters after the fused Adam step, equal the single-GPU result on the
global batch to rel-L2 <= 1e-5 per tensor - auxiliary task off and on, even and ragged shards (B = 7 over 2 ranks),
global counts handed in or exchanged; and the `train()` driver under 2 ranks (row f4) ends with the parameters a
single rank reaches on the same global batches.
"""
import json
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-5


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _need_two_gpus():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")


def _model_and_batch(aux, B, dev):
    import multimodal_seq2seq_gscan_b200 as pkg
    from oracle import gscan_oracle as O
    cfg = dict(O.CONFIGS["comp"])
    cfg["auxiliary_task"] = aux
    params = O.synthetic_params(cfg, 31, scale=2.0)
    batch = O.synthetic_batch(cfg, batch_size=B, seed=32, max_tgt_len=14)
    model = pkg.Model(**O.model_kwargs(cfg)).to(dev)
    from multimodal_seq2seq_gscan_b200.synthetic import full_state_dict
    model.load_state_dict(full_state_dict(params), strict=True)
    return cfg, model, batch


def _step(trainer, batch, lo, hi, dev, aux, global_counts=None, cut=True):
    cmd_len, tgt_len = batch["cmd_lengths"][lo:hi], batch["tgt_lengths"][lo:hi]
    Ti = int(cmd_len.max())
    Tt = batch["targets"].shape[1] if (aux or not cut) else int(tgt_len.max())
    c = torch.tensor(batch["commands"][lo:hi, :Ti], device=dev)
    s = torch.tensor(batch["situations"][lo:hi], device=dev)
    t = torch.tensor(batch["targets"][lo:hi, :Tt], device=dev)
    pos = torch.tensor(batch["target_positions"][lo:hi], device=dev) if aux else None
    return trainer.train_step(c, cmd_len, s, t, tgt_len, pos, global_counts=global_counts)


def _trainer_worker(rank, world, port, aux, B, know_global, out_dir):
    import torch.distributed as dist
    from multimodal_seq2seq_gscan_b200 import dp
    from multimodal_seq2seq_gscan_b200.trainer import FusedTrainer
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        cfg, model, batch = _model_and_batch(aux, B, dev)
        trainer = FusedTrainer(model, distributed=True)
        lo, hi = dp.shard_bounds(B, rank, world)
        gc = (float(np.sum(batch["tgt_lengths"] - 1)), float(B)) if know_global else None
        losses = []
        for _ in range(2):     # two steps: the second one starts from all-reduced, Adam-updated parameters
            losses.append(_step(trainer, batch, lo, hi, dev, aux, gc))
            grad = trainer.last_flat_grad
            n = trainer._n
        total = torch.stack(losses)
        dist.all_reduce(total)
        torch.save({"grad": (grad[:n] / grad[n]).cpu(), "counts": grad[n:n + 2].cpu(),
                    "param": trainer.flat_param.cpu(), "loss": total.cpu()}, os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("aux,know_global", [(False, False), (True, True), (True, False)])
@pytest.mark.parametrize("B", [7, 16])
def test_two_ranks_equal_one_gpu(aux, know_global, B, tmp_path):
    _need_two_gpus()
    import torch.multiprocessing as mp
    from multimodal_seq2seq_gscan_b200.trainer import FusedTrainer
    mp.spawn(_trainer_worker, args=(2, _free_port(), aux, B, know_global, str(tmp_path)), nprocs=2, join=True)
    outs = [torch.load(os.path.join(tmp_path, f"r{r}.pt"), weights_only=False) for r in range(2)]
    # one GPU, the concatenated batch, same two steps
    dev = torch.device("cuda:0")
    cfg, model, batch = _model_and_batch(aux, B, dev)
    ref = FusedTrainer(model, distributed=False)
    losses = [_step(ref, batch, 0, B, dev, aux, cut=False) for _ in range(2)]
    n = ref._n
    g_ref = ref.last_flat_grad[:n].cpu()
    p_ref = ref.flat_param.cpu()
    sizes, offsets = ref._sizes, ref._offsets
    n_tok = float(np.sum(batch["tgt_lengths"] - 1))
    for r, o in enumerate(outs):
        assert o["counts"].tolist() == [n_tok, float(B)]
        for i, (sz, off) in enumerate(zip(sizes, offsets[:-1])):
            if sz == 0:
                continue
            a, b = o["grad"][int(off):int(off) + sz].double(), g_ref[int(off):int(off) + sz].double()
            assert (a - b).norm() <= TOL * max(b.norm(), 1e-12), (r, "grad", i, ((a - b).norm() / b.norm()).item())
            a, b = o["param"][int(off):int(off) + sz].double(), p_ref[int(off):int(off) + sz].double()
            assert (a - b).norm() <= TOL * max(b.norm(), 1e-12), (r, "param", i)
        # the loss shares of the ranks add up to the global-batch loss, step by step
        for k in range(2):
            assert abs(o["loss"][k].item() - losses[k].item()) <= 1e-5 * max(1.0, abs(losses[k].item()))
    assert torch.equal(outs[0]["param"], outs[1]["param"])      # identical replicas


def _train_args(tmp, **over):
    a = dict(data_path=os.path.join(tmp, "dataset.txt"), data_directory=tmp, generate_vocabularies=False,
             input_vocab_path="training_input_vocab.txt", target_vocab_path="training_target_vocab.txt",
             embedding_dimension=25, num_encoder_layers=1, encoder_dropout_p=0.0, encoder_bidirectional=True,
             training_batch_size=6, test_batch_size=5, max_decoding_steps=30, num_decoder_layers=1,
             decoder_dropout_p=0.0, cnn_kernel_size=7, cnn_dropout_p=0.0, cnn_hidden_num_channels=50,
             simple_situation_representation=True, decoder_hidden_size=100, encoder_hidden_size=100,
             learning_rate=2e-3, adam_beta_1=0.9, adam_beta_2=0.999, lr_decay=0.9, lr_decay_steps=20000,
             resume_from_file="", max_training_iterations=9, output_directory=tmp, print_every=100,
             evaluate_every=100, conditional_attention=True, auxiliary_task=True, weight_target_loss=0.3,
             attention_type="bahdanau", k=0, max_training_examples=None, seed=3, max_testing_examples=None)
    a.update(over)
    return a


def _train_worker(rank, world, port, tmp):
    import torch.distributed as dist
    from multimodal_seq2seq_gscan_b200 import train as T
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        out = T.train(**_train_args(tmp))
        torch.save({"param": out["trainer"].flat_param.cpu(), "iterations": out["iterations"]},
                   os.path.join(tmp, f"train_r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_train_driver_two_ranks_matches_one(tmp_path):
    """Row f4: `train()` under 2 ranks.  23 training examples in global batches of 6 -> 6, 6, 6, 5 (ragged shards
    3 + 2) per epoch; with dropout off the run is deterministic and must end where a single rank ends."""
    _need_two_gpus()
    import torch.multiprocessing as mp
    from multimodal_seq2seq_gscan_b200 import train as T
    from multimodal_seq2seq_gscan_b200.dataset import GroundedScanDataset
    tmp = str(tmp_path)
    data = json.load(open(os.path.join(GOLD, "dataset_small.txt")))
    data["examples"]["dev"] = data["examples"].pop("test")
    json.dump(data, open(os.path.join(tmp, "dataset.txt"), "w"))
    # vocabularies written once, up front, so that both runs read the same files
    ds = GroundedScanDataset(os.path.join(tmp, "dataset.txt"), tmp, split="train",
                             input_vocabulary_file="training_input_vocab.txt",
                             target_vocabulary_file="training_target_vocab.txt", generate_vocabulary=True)
    ds.save_vocabularies("training_input_vocab.txt", "training_target_vocab.txt")
    mp.spawn(_train_worker, args=(2, _free_port(), tmp), nprocs=2, join=True)
    outs = [torch.load(os.path.join(tmp, f"train_r{r}.pt"), weights_only=False) for r in range(2)]
    one = T.train(**_train_args(tmp))
    p_ref = one["trainer"].flat_param.cpu().double()
    assert outs[0]["iterations"] == outs[1]["iterations"] == one["iterations"]
    assert torch.equal(outs[0]["param"], outs[1]["param"])
    rel = ((outs[0]["param"].double() - p_ref).norm() / p_ref.norm()).item()
    assert rel <= 1e-5, rel
