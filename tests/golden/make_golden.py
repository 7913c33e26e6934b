"""Generate golden vectors from the UNMODIFIED reference (run in the build container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Imports ``seq2seq`` from /root/reference (with a ``GroundedScan`` stub because ``gym`` is not
installed - SURVEY.md 8(c)), loads deterministic numpy-generated parameters into the reference
``Model``, runs its forward / loss / backward and its batch-size-1 ``predict`` loop on CPU in
float32 and float64, and stores the OUTPUTS as small .npz fixtures.  Inputs and parameters are
not stored: they are re-created from the seed by ``oracle.gscan_oracle.synthetic_params`` /
``synthetic_batch`` (numpy PCG64, machine independent).  The GPU box never needs /root/reference.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

from oracle import gscan_oracle as O  # noqa: E402


def import_reference():
    stub = types.ModuleType("GroundedScan")
    stub_ds = types.ModuleType("GroundedScan.dataset")

    class GroundedScan:  # noqa: D401 - placeholder for the gym-dependent generator
        pass

    stub_ds.GroundedScan = GroundedScan
    stub.dataset = stub_ds
    sys.modules.setdefault("GroundedScan", stub)
    sys.modules.setdefault("GroundedScan.dataset", stub_ds)
    sys.path.insert(0, "/root/reference")
    from seq2seq.model import Model  # noqa: E402
    from seq2seq.predict import predict  # noqa: E402
    from seq2seq.helpers import sequence_accuracy  # noqa: E402
    return Model, predict, sequence_accuracy


def load_params(model, params):
    sd = {k: v.clone() for k, v in params.items()}
    for att in ("textual_attention", "visual_attention"):
        for layer in ("key_layer", "query_layer", "energy_layer"):
            sd[f"attention_decoder.{att}.{layer}.weight"] = sd[f"{att}.{layer}.weight"]
    model.load_state_dict(sd, strict=True)


CASES = [
    # name, config, overrides, batch kwargs, dtype
    ("tiny_aux", "tiny", {}, dict(batch_size=5, max_cmd_len=6, min_cmd_len=3, max_tgt_len=9, min_tgt_len=3), "float64"),
    ("tiny_nocond", "tiny", {"conditional_attention": False, "auxiliary_task": False},
     dict(batch_size=4, max_cmd_len=5, min_cmd_len=3, max_tgt_len=6, min_tgt_len=3), "float64"),
    ("demo", "demo", {}, dict(batch_size=5, max_cmd_len=8, min_cmd_len=4, max_tgt_len=14, min_tgt_len=3), "float64"),
    ("demo_f32", "demo", {}, dict(batch_size=5, max_cmd_len=8, min_cmd_len=4, max_tgt_len=14, min_tgt_len=3), "float32"),
    ("comp_small", "comp", {}, dict(batch_size=3, max_cmd_len=10, min_cmd_len=5, max_tgt_len=12, min_tgt_len=3), "float64"),
    ("comp_aux", "comp", {"auxiliary_task": True}, dict(batch_size=6, max_cmd_len=10, min_cmd_len=5, max_tgt_len=30, min_tgt_len=3), "float64"),
    ("tlen_small", "tlen", {}, dict(batch_size=2, max_cmd_len=9, min_cmd_len=5, max_tgt_len=20, min_tgt_len=17), "float64"),
]

WEIGHT_TARGET_LOSS = 0.3
SEED = 1234


def run_case(Model, predict, sequence_accuracy, name, cfg_name, overrides, batch_kw, dtype_name):
    cfg = dict(O.CONFIGS[cfg_name])
    cfg.update(overrides)
    dtype = getattr(torch, dtype_name)
    params = O.synthetic_params(cfg, SEED, scale=2.0, dtype=dtype)
    batch = O.synthetic_batch(cfg, seed=SEED + 1, **batch_kw)
    model = Model(**O.model_kwargs(cfg))
    if dtype == torch.float64:
        model = model.double()
    load_params(model, params)
    model.train()  # dropout probabilities are 0 in these configs
    commands = torch.tensor(batch["commands"])
    situations = torch.tensor(batch["situations"], dtype=dtype)
    targets = torch.tensor(batch["targets"])
    positions = torch.tensor(batch["target_positions"])
    logp, aux = model(commands_input=commands, commands_lengths=batch["cmd_lengths"],
                      situations_input=situations, target_batch=targets,
                      target_lengths=batch["tgt_lengths"])
    loss = model.get_loss(logp, targets)
    out = {"logp": logp.detach().numpy(), "nll": loss.detach().numpy()}
    if cfg["auxiliary_task"]:
        aux_loss = model.get_auxiliary_loss(aux, positions)
        out["aux_logp"] = aux.detach().numpy()
        out["aux_nll"] = aux_loss.detach().numpy()
        loss = loss + WEIGHT_TARGET_LOSS * aux_loss
        out["aux_accuracy"] = np.float64(model.get_auxiliary_accuracy(aux, positions))
    out["loss"] = loss.detach().numpy()
    acc, exact = model.get_metrics(logp, targets)
    out["accuracy"] = np.float64(acc)
    out["exact_match"] = np.float64(exact)
    loss.backward()
    named = dict(model.named_parameters())
    for pname, _ in O.param_shapes(cfg):
        out["grad." + pname] = named[pname].grad.detach().numpy()

    # encode_input / sub-module outputs (eval mode) for per-kernel parity
    model.eval()
    with torch.no_grad():
        enc = model.encode_input(commands_input=commands, commands_lengths=batch["cmd_lengths"],
                                 situations_input=situations)
        out["encoded_situations"] = enc["encoded_situations"].numpy()
        out["encoder_outputs"] = enc["encoded_commands"]["encoder_outputs"].numpy()
        out["hidden_states"] = enc["hidden_states"].numpy()

    # greedy decode through the reference's own predict() loop, batch size 1
    def iterator():
        for b in range(commands.shape[0]):
            n_in = int(batch["cmd_lengths"][b])
            n_tg = int(batch["tgt_lengths"][b])
            yield (commands[b:b + 1, :n_in], [n_in], [""], situations[b:b + 1], [{}],
                   targets[b:b + 1, :n_tg], [n_tg], torch.zeros(1, dtype=torch.long),
                   positions[b:b + 1])

    max_steps = 12
    seqs, accs, aux_accs, betasum = [], [], [], []
    for (_inp, _d, _s, output_sequence, target_sequence, att_cmd, att_sit, aux_acc) in predict(
            iterator(), model=model, max_decoding_steps=max_steps, pad_idx=0, sos_idx=1, eos_idx=2):
        seqs.append(output_sequence)
        accs.append(sequence_accuracy(output_sequence, target_sequence[0].tolist()[1:-1]))
        aux_accs.append(float(aux_acc))
    L = max(1, max(len(s) for s in seqs))
    arr = -np.ones((len(seqs), L), dtype=np.int64)
    for i, s in enumerate(seqs):
        arr[i, :len(s)] = s
    out["greedy_sequences"] = arr
    out["greedy_lengths"] = np.array([len(s) for s in seqs], dtype=np.int64)
    out["greedy_accuracy"] = np.array(accs)
    out["greedy_aux_accuracy"] = np.array(aux_accs)
    out["greedy_max_steps"] = np.int64(max_steps)

    # a variant where EOS can never be produced (eos_idx = -1): exercises the N+1 cap
    # (``<=`` at predict.py:101)
    seqs2 = []
    for (_inp, _d, _s, output_sequence, *_rest) in predict(
            iterator(), model=model, max_decoding_steps=5, pad_idx=0, sos_idx=1, eos_idx=-1):
        seqs2.append(output_sequence)
    out["greedy_noeos_sequences"] = np.array(seqs2, dtype=np.int64)   # all have length 6 = N+1
    return cfg, batch_kw, out


def run_greedy_long(Model, predict):
    """Round 2: the reference's own ``predict()`` at the FULL decoding length (max_decoding_steps = 120, the value of
    all_experiments.sh) on the compositional shape with the auxiliary task, float32 as the reference runs it, with
    the per-step attention weights it returns (predict.py:108-109) - once with the real EOS and once with EOS
    unreachable (every sequence runs the full 121 steps: the N+1 cap of predict.py:101)."""
    cfg = dict(O.CONFIGS["comp"])
    cfg["auxiliary_task"] = True
    batch_kw = dict(batch_size=4, max_cmd_len=10, min_cmd_len=5, max_tgt_len=12, min_tgt_len=3)
    # (seed and scale picked so that the four sequences end after 13, 3, 0 and - never - 121 tokens)
    params = O.synthetic_params(cfg, SEED + 13, scale=3.0, dtype=torch.float32)
    batch = O.synthetic_batch(cfg, seed=SEED + 14, **batch_kw)
    model = Model(**O.model_kwargs(cfg))
    load_params(model, params)
    model.eval()
    commands = torch.tensor(batch["commands"])
    situations = torch.tensor(batch["situations"])
    targets = torch.tensor(batch["targets"])
    positions = torch.tensor(batch["target_positions"])
    B, Ti, M = commands.shape[0], commands.shape[1], cfg["grid_size"] ** 2

    def iterator():
        for b in range(B):
            n_in, n_tg = int(batch["cmd_lengths"][b]), int(batch["tgt_lengths"][b])
            yield (commands[b:b + 1, :n_in], [n_in], [""], situations[b:b + 1], [{}], targets[b:b + 1, :n_tg], [n_tg],
                   torch.zeros(1, dtype=torch.long), positions[b:b + 1])

    out = {}
    max_steps = 120
    for tag, eos in (("eos", 2), ("noeos", -1)):
        T = max_steps + 1
        seq = -np.ones((B, T), dtype=np.int64)
        lens = np.zeros(B, dtype=np.int64)
        alphas = np.zeros((B, T, Ti), dtype=np.float32)
        betas = np.zeros((B, T, M), dtype=np.float32)
        aux_acc = np.zeros(B)
        with torch.no_grad():
            for b, (_i, _d, _s, output_sequence, _t, att_cmd, att_sit, acc) in enumerate(
                    predict(iterator(), model=model, max_decoding_steps=max_steps, pad_idx=0, sos_idx=1, eos_idx=eos)):
                n = len(output_sequence)
                seq[b, :n] = output_sequence
                lens[b] = n
                assert len(att_cmd) == n and len(att_sit) == n
                for t in range(n):
                    a = np.asarray(att_cmd[t], dtype=np.float32).reshape(-1)      # [n_in] of this example
                    alphas[b, t, :a.size] = a
                    betas[b, t] = np.asarray(att_sit[t], dtype=np.float32).reshape(-1)
                aux_acc[b] = float(acc)
        out[f"{tag}_sequences"], out[f"{tag}_lengths"] = seq, lens
        out[f"{tag}_alphas"], out[f"{tag}_betas"], out[f"{tag}_aux_accuracy"] = alphas, betas, aux_acc
    meta = dict(case_name="greedy_long", cfg_name="comp", overrides=repr({"auxiliary_task": True}),
                batch_kw=repr(batch_kw), dtype="float32", seed=SEED + 13, weight_target_loss=WEIGHT_TARGET_LOSS,
                param_scale=3.0, max_decoding_steps=max_steps)
    path = os.path.join(HERE, "greedy_long.npz")
    np.savez_compressed(path, __meta__=np.array(repr(meta)), **out)
    print(f"greedy_long: wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB), lengths with EOS "
          f"{out['eos_lengths'].tolist()}, without {out['noeos_lengths'].tolist()}")


def main():
    Model, predict, sequence_accuracy = import_reference()
    torch.manual_seed(0)
    if "--only-greedy-long" in sys.argv:
        run_greedy_long(Model, predict)
        return
    for case in CASES:
        name = case[0]
        cfg, batch_kw, out = run_case(Model, predict, sequence_accuracy, *case)
        meta = dict(case_name=name, cfg_name=case[1], overrides=repr(case[2]), batch_kw=repr(batch_kw),
                    dtype=case[4], seed=SEED, weight_target_loss=WEIGHT_TARGET_LOSS, param_scale=2.0)
        # full-size gradients are stored as float32 to keep fixtures small
        store = {}
        for k, v in out.items():
            v = np.asarray(v)
            if v.dtype == np.float64 and v.size > 4096:
                v = v.astype(np.float32)
            store[k] = v
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, __meta__=np.array(repr(meta)), **store)
        print(f"{name}: wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB), loss={float(out['loss']):.6f}, "
              f"greedy_len={out['greedy_lengths'].tolist()}")
    run_greedy_long(Model, predict)


if __name__ == "__main__":
    main()
