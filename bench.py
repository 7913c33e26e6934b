#!/usr/bin/env python
"""Benchmark of the gSCAN seq2seq hot path on B200 (BASELINE.json metric: train examples/sec at
batch 200 per GPU, plus batched greedy-decode sequences/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full training iteration of the compositional_splits shape (BASELINE.json configs[1]:
grid 6, B=200 per GPU, k=7, E=25, H=100, Ti=10, Tt=121 - all padded steps computed as the reference
does, paper dropout 0.3/0.3/0.1, no auxiliary task): forward + NLL loss + backward + (ONE all-reduce of
gradients and loss normalisers when N>1) + Adam.  Rank 0 prints ONE JSON line.

  value     examples/s with the batch already resident in HBM; every step is bracketed by its own
            CUDA events on the launching stream and L2 is flushed (256 MiB memset) between steps;
            the per-rank totals are reduced with MAX over ranks.
  e2e       the same step driven from pinned HOST buffers through the public API
            (Model/FusedTrainer): per step the H2D copies of the batch and an async D2H copy of the
            loss, which the host reads one step later; wall clock over >= 100 consecutive steps with a
            device synchronize at both ends.
  roofline  the dominant kernel (decoder backward cluster sweep) timed live with CUDA events recorded by
            the library on the same stream (gscan_profile).  The sweep is a latency-bound recurrence (121
            dependent steps): the object carries the (tiny, honest) fraction of the measured tensor peak, and
            the figure that means something - microseconds per decoder step against a per-phase latency model
            of the critical path (profiles/r02_ncu_sweeps.md) - plus the ncu numbers of the shipping kernels.
  decode    batched greedy decoding (configs[3]): device-timed value, e2e from pinned host inputs with the
            tokens and lengths copied back, us per decoding step, the reference's batch-1 predict() timed on
            the host cores beside it; with --gpus N every rank decodes its own replica batch.
  cpu_baseline / --impl reference
            the UNMODIFIED reference (oracle/_ref, vendored by oracle/make_ref.py; kind "reference") on all
            host cores, same shape, same step definition, CUDA hidden from it; falls back to the CPU port
            (oracle/gscan_oracle.py; kind "port") when the vendored copy is absent.
  reference_gpu_eager
            the same unmodified reference in PyTorch eager mode on the B200 itself (TF32 off and on):
            the same-box GPU bar of SURVEY.md 8(d).
"""
from __future__ import annotations

import os
import sys

# the reference arm and its helper legs must not see the GPU: the reference picks its device at import time
if "--impl=reference" in sys.argv or ("--impl" in sys.argv and sys.argv[sys.argv.index("--impl") + 1:][:1] == ["reference"]):
    os.environ["CUDA_VISIBLE_DEVICES"] = ""

import argparse
import gc
import json
import statistics
import subprocess
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1] - the configuration the metric is quoted on (default)
    "comp": "compositional_splits: G=6 C=16 F=50 k3=7 E=25 H=100 Vi=21 V=9 B=200/GPU Ti=10 Tt=121 (all padded steps), no aux, dropout .3/.3/.1",
    # configs[2]: same with the auxiliary target-position task (weight 0.3)
    "comp_aux": "compositional_splits + auxiliary target-position task (weight_target_loss 0.3): G=6 C=16 F=50 k3=7 E=25 H=100 Vi=21 V=9 B=200/GPU Ti=10 Tt=121, dropout .3/.3/.1",
    # configs[4]: target_length_split shape (long action sequences, 13x13 third convolution)
    "tlen": "target_length_split: G=6 C=16 F=50 k3=13 E=25 H=100 Vi=17 V=8 B=200/GPU Ti=10 Tt=121 (target lengths 17..121), no aux, dropout .3/.3/.1",
}
SEED = 1234
B_PER_GPU = 200
# algorithmic FLOPs (2 per MAC) per example per decoder step inside the recurrent sweeps:
# forward: the four dependent mat-vec stages (600+500+100+400 rows x 100) + both attentions (46 keys x 100 x 2)
SWEEP_FWD_FLOP = 2 * (1600 * 100 + 46 * 100 * 2)
# backward: the transposed mat-vecs (same 160k MAC) + attention backward (46 keys x 100 x 4)
SWEEP_BWD_FLOP = 2 * (1600 * 100 + 46 * 100 * 4)
# whole training step per example: 3 x (7.62M + 121 x 0.506M)
STEP_FLOP_PER_EXAMPLE = 206.5e6
# greedy decoding per example: 7.62M + 0.506M per generated step
DECODE_FLOP_FIXED, DECODE_FLOP_PER_STEP = 7.62e6, 0.506e6


def bench_cfg(workload: str):
    from multimodal_seq2seq_gscan_b200 import synthetic
    cfg = dict(synthetic.CONFIGS["tlen" if workload == "tlen" else "comp"])
    cfg.update(encoder_dropout_p=0.3, decoder_dropout_p=0.3, cnn_dropout_p=0.1, auxiliary_task=(workload == "comp_aux"))
    return cfg


def bench_config(workload: str, world: int) -> dict:
    """The `config` object of the JSON line - the SAME dict for both arms (the driver compares them)."""
    return {"workload": WORKLOADS[workload], "global_batch": world * B_PER_GPU,
            "parallelism": f"dp{world}" if world > 1 else "single",
            "step": "forward + NLL + backward + (one all-reduce of gradients and loss normalisers when N > 1) + Adam",
            "l2": "flushed between timed steps (256 MiB memset); the step's own workspace (~300 MB) also exceeds L2"}


def read_peaks():
    """Measured peaks written by the driver (MEASURED_PEAKS.json: HBM GB/s and dense bf16 TFLOP/s of this pool's
    B200s; a value may be a number or a {burst, sustained} object - the sweep is timed inside a long step, so the
    sustained figure applies), else the fallback of B200_PROFILING.md."""
    fallback = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "sm_max_mhz": 1965.0, "source": "fallback"}
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if not os.path.exists(path):
        return fallback

    def number(v):
        if isinstance(v, dict):
            for key in ("sustained", "sustained_tflops", "sustained_gbs", "value", "burst"):
                if key in v:
                    return float(v[key])
            return float(next(iter(v.values())))
        return float(v)

    try:
        with open(path) as f:
            p = json.load(f)
        out = dict(fallback, source="measured")
        for key in ("hbm_gbs", "bf16_tflops", "sm_max_mhz"):
            for cand in (key + "_sustained", key):
                if cand in p:
                    out[key] = number(p[cand])
                    break
        return out
    except Exception as exc:      # a malformed file must not take the bench line down
        return dict(fallback, source=f"fallback (MEASURED_PEAKS.json unreadable: {type(exc).__name__})")


def read_ncu_profile():
    """ncu --set full numbers of the SHIPPING sweeps and the per-phase latency model (profiles/, regenerated
    each round by tools/ncu_sweeps_summary.py from the capture of the same bench command)."""
    for name in ("r02_ncu_sweeps.json", "r01_v3_ncu_sweeps.json"):
        path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(path):
            with open(path) as f:
                d = json.load(f)
            d["file"] = "profiles/" + name
            return d
    return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu_index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            parts = [x.strip() for x in row.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_host_batch(cfg, seed, workload):
    from multimodal_seq2seq_gscan_b200 import synthetic
    return synthetic.synthetic_batch(cfg, batch_size=B_PER_GPU, seed=seed, min_tgt_len=17 if workload == "tlen" else 3)


# --------------------------------------------------------------------------------------------
# reference arm: the UNMODIFIED reference on the host cores (or on the GPU for the eager side leg).
# The only places bench.py executes oracle/.
# --------------------------------------------------------------------------------------------
def _reference_model(cfg, device):
    """(kind, model-or-params, step function).  The vendored reference when present, else the CPU port."""
    from oracle import gscan_oracle as O
    from oracle import ref_loader
    params = O.synthetic_params(cfg, SEED)
    if ref_loader.available():
        ref = ref_loader.load()
        model = ref.model.Model(**O.model_kwargs(cfg))
        sd = {k: v.clone() for k, v in params.items()}
        for att in ("textual_attention", "visual_attention"):
            for layer in ("key_layer", "query_layer", "energy_layer"):
                sd[f"attention_decoder.{att}.{layer}.weight"] = sd[f"{att}.{layer}.weight"]
        model.load_state_dict(sd, strict=True)
        return "reference", model.to(device), ref
    return "port", {k: v.clone().to(device).requires_grad_(True) for k, v in params.items()}, None


def reference_train_steps(cfg, workload, steps, warmup, device="cpu", threads=None):
    """forward + loss + backward + Adam exactly as the reference's training loop does it (train.py:96-113: train
    mode, paper dropout, torch.optim.Adam), on `device`.  Returns (examples/s from the median step, seconds per
    step, threads, kind)."""
    from oracle import gscan_oracle as O
    threads = threads or len(os.sched_getaffinity(0))
    torch.set_num_threads(threads)
    dev = torch.device(device)
    kind, model, ref = _reference_model(cfg, dev)
    batch = make_host_batch(cfg, SEED + 1, workload)
    commands, targets = torch.tensor(batch["commands"], device=dev), torch.tensor(batch["targets"], device=dev)
    situations = torch.tensor(batch["situations"], device=dev)
    positions = torch.tensor(batch["target_positions"], device=dev)
    B, Ti, Tt = commands.shape[0], commands.shape[1], targets.shape[1]
    M, D = cfg["grid_size"] ** 2, 3 * cfg["cnn_hidden_num_channels"]
    E, H = cfg["embedding_dimension"], cfg["decoder_hidden_size"]
    if kind == "reference":
        model.train()
        opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-3)
    else:
        opt = torch.optim.Adam(list(model.values()), lr=1e-3)

    def mask(shape, p):
        return None if p <= 0 else torch.empty(shape, device=dev).bernoulli_(1 - p) / (1 - p)

    times = []
    for it in range(warmup + steps):
        if dev.type == "cuda":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        if kind == "reference":
            logp, aux = model(commands_input=commands, commands_lengths=batch["cmd_lengths"],
                              situations_input=situations, target_batch=targets, target_lengths=batch["tgt_lengths"])
            loss = model.get_loss(logp, targets)
            if cfg["auxiliary_task"]:
                loss = loss + 0.3 * model.get_auxiliary_loss(aux, positions)
        else:
            drop = {"cnn": mask((B, M, D), cfg["cnn_dropout_p"]), "enc": mask((B, Ti, E), cfg["encoder_dropout_p"]),
                    "dec": mask((B, Tt, H), cfg["decoder_dropout_p"])}
            logp, aux = O.model_forward(model, commands, batch["cmd_lengths"], situations, targets,
                                        cfg["conditional_attention"], cfg["auxiliary_task"], dropout=drop)
            loss = O.nll_loss(logp, targets)
            if cfg["auxiliary_task"]:
                loss = loss + 0.3 * O.aux_nll_loss(aux, positions)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        if dev.type == "cuda":
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    med = statistics.median(times)
    return B / med, times, threads, kind


def reference_decode(cfg, workload, n_seq, device="cpu", threads=None):
    """The reference's own greedy loop at batch size 1 (predict.py:57-128) over `n_seq` sequences with EOS
    unreachable (121 steps each - the worst case the GPU number is quoted on).  Returns (sequences/s, kind)."""
    from oracle import gscan_oracle as O
    threads = threads or len(os.sched_getaffinity(0))
    torch.set_num_threads(threads)
    dev = torch.device(device)
    kind, model, ref = _reference_model(cfg, dev)
    batch = make_host_batch(cfg, SEED + 1, workload)
    commands, situations = torch.tensor(batch["commands"], device=dev), torch.tensor(batch["situations"], device=dev)
    targets, positions = torch.tensor(batch["targets"], device=dev), torch.tensor(batch["target_positions"], device=dev)

    def iterator(n):
        for b in range(n):
            n_in, n_tg = int(batch["cmd_lengths"][b]), int(batch["tgt_lengths"][b])
            yield (commands[b:b + 1, :n_in], [n_in], [""], situations[b:b + 1], [{}], targets[b:b + 1, :n_tg], [n_tg],
                   torch.zeros(1, dtype=torch.long, device=dev), positions[b:b + 1])

    def run(n):
        with torch.no_grad():
            if kind == "reference":
                model.eval()
                for _ in ref.predict.predict(iterator(n), model=model, max_decoding_steps=120, pad_idx=0, sos_idx=1,
                                             eos_idx=-1):
                    pass
            else:
                for b in range(n):
                    n_in = int(batch["cmd_lengths"][b])
                    O.greedy_decode(model, commands[b:b + 1, :n_in], batch["cmd_lengths"][b:b + 1],
                                    situations[b:b + 1], 120, eos_idx=-1)
        if dev.type == "cuda":
            torch.cuda.synchronize()

    run(1)
    t0 = time.perf_counter()
    run(n_seq)
    return n_seq / (time.perf_counter() - t0), kind, threads


def run_reference_arm(args, json_out):
    """`--impl reference`: rank 0 alone; CUDA is hidden from this process (see the top of the file)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cfg = bench_cfg(args.workload)
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    value, times, threads, kind = reference_train_steps(cfg, args.workload, steps, warmup)
    sample = (f"{steps} full steps (B=200, Tt=121, train mode, fwd+loss+bwd+Adam) after {warmup} warm-up, median "
              f"{statistics.median(times):.3f} s, min {min(times):.3f} s")
    what = ("the unmodified reference (oracle/_ref, vendored from /root/reference by oracle/make_ref.py)" if kind == "reference"
            else "the CPU port oracle/gscan_oracle.py (oracle/_ref not present)")
    line = {
        "impl": "reference", "metric": "train_examples_per_sec", "value": value, "unit": "examples/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * statistics.median(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args.workload, args.gpus),
        "device": f"host CPU, {threads} threads; {what}; one process regardless of --gpus",
        "cpu_baseline": {"value": value, "unit": "examples/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "examples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if args.with_decode:
        sps, kind_d, _ = reference_decode(cfg, args.workload, args.decode_seqs)
        line["decode"] = {"seqs_per_sec": sps, "steps_per_sec": sps * 121, "batch": 1, "kind": kind_d, "cores": threads,
                          "sample": f"{args.decode_seqs} sequences at batch size 1 through predict() (predict.py:57-128), "
                                    "EOS unreachable: 121 steps each"}
    print(json.dumps(line), file=json_out, flush=True)


def run_reference_gpu_eager(args, json_out):
    """Hidden side leg (`--impl reference-gpu`): the unmodified reference in PyTorch eager mode ON the B200."""
    cfg = bench_cfg(args.workload)
    out = {}
    for tf32 in (False, True):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        value, times, _, kind = reference_train_steps(cfg, args.workload, max(1, args.steps), max(1, args.warmup), "cuda")
        out["tf32_on" if tf32 else "tf32_off"] = {"examples_per_sec": value, "ms_per_step": 1e3 * statistics.median(times)}
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    sps, kind_d, _ = reference_decode(cfg, args.workload, 8, "cuda")
    out["decode_batch1_seqs_per_sec"] = sps
    out["kind"] = kind
    out["sample"] = (f"{args.steps} steps after {args.warmup} warm-up per TF32 setting, wall clock with a device "
                     "synchronize around every step; decode: 8 sequences x 121 steps at batch size 1 (TF32 off)")
    print(json.dumps(out), file=json_out, flush=True)


def _side_leg(argv, timeout, hide_cuda):
    """Run another leg of this file in a subprocess and parse its single JSON line (None on failure)."""
    env = dict(os.environ)
    if hide_cuda:
        env["CUDA_VISIBLE_DEVICES"] = ""
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    try:
        res = subprocess.run([sys.executable, os.path.abspath(__file__)] + argv, env=env, capture_output=True, text=True,
                             timeout=timeout)
        lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
        return json.loads(lines[-1]) if lines else {"error": (res.stderr or "no output")[-300:]}
    except Exception as exc:
        return {"error": f"{type(exc).__name__}: {exc}"[:300]}


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
def _claim_stdout():
    """Libraries under us print to fd 1 (NCCL's version banner when NCCL_DEBUG=VERSION, for one); the driver
    expects exactly ONE JSON line there.  Keep a private copy of the real stdout for that line and point
    fd 1 at stderr for everything else."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


class Workload:
    """Model + trainer + one synthetic batch (pinned on the host and resident on the device) of one workload."""

    def __init__(self, key, dev, rank, world, distributed):
        import multimodal_seq2seq_gscan_b200 as pkg
        from multimodal_seq2seq_gscan_b200 import synthetic as S
        from multimodal_seq2seq_gscan_b200.trainer import FusedTrainer
        self.key, self.dev, self.world, self.distributed = key, dev, world, distributed
        self.cfg = cfg = bench_cfg(key)
        self.model = pkg.Model(**S.model_kwargs(cfg)).to(dev)
        self.model.load_state_dict(S.full_state_dict(S.synthetic_params(cfg, SEED)), strict=True)
        self.trainer = FusedTrainer(self.model, distributed=distributed)
        self.host = host = make_host_batch(cfg, SEED + 1 + rank, key)
        self.pinned = {k: torch.from_numpy(np.ascontiguousarray(host[k])).pin_memory()
                       for k in ("commands", "situations", "targets")}
        self.resident = {k: v.to(dev) for k, v in self.pinned.items()}
        self.cmd_len, self.tgt_len = host["cmd_lengths"], host["tgt_lengths"]
        self.positions = torch.from_numpy(host["target_positions"]).to(dev) if cfg["auxiliary_task"] else None
        # The loss normalisers of the GLOBAL batch: every rank can compute them from the rank numbers alone (the
        # synthetic shard of rank r is seeded with r), as train.py can from the global batch it holds.  Needed
        # before the backward pass only by the auxiliary loss; without it the counts travel with the gradients.
        self.global_counts = None
        if distributed and cfg["auxiliary_task"]:
            n_tok = sum(float(np.sum(make_host_batch(cfg, SEED + 1 + r, key)["tgt_lengths"] - 1)) for r in range(world))
            self.global_counts = (n_tok, float(world * B_PER_GPU))

    def step_resident(self):
        r = self.resident
        return self.trainer.train_step(r["commands"], self.cmd_len, r["situations"], r["targets"], self.tgt_len,
                                       self.positions, global_counts=self.global_counts)

    def step_from_host(self):
        p, dev = self.pinned, self.dev
        c = p["commands"].to(dev, non_blocking=True)
        s = p["situations"].to(dev, non_blocking=True)
        t = p["targets"].to(dev, non_blocking=True)
        return self.trainer.train_step(c, self.cmd_len, s, t, self.tgt_len, self.positions,
                                       global_counts=self.global_counts)


def measure_training(wl, steps, warmup, flush, barrier, dist, lib):
    """(ms per step [max over ranks of per-step CUDA-event times], launches) with the batch resident.
    The Python garbage collector is run before the barrier and switched off inside the timed loop (a generation-2 pass is
    1-2 ms of host time, enough to drain the launch queue once in a few dozen steps)."""
    for _ in range(warmup):
        wl.step_resident()
    barrier()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    launches0 = lib.gscan_launch_count()
    if wl.distributed:
        wl.trainer.collective_events = []
    gc.collect()
    gc.disable()          # no collector pauses inside the timed region (a generation-2 pass is 1-2 ms of host time)
    barrier()
    try:
        for i in range(steps):
            flush.zero_()
            starts[i].record()
            wl.step_resident()
            ends[i].record()
    finally:
        gc.enable()
    barrier()
    launches = lib.gscan_launch_count() - launches0
    per_step = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    if os.environ.get("GSCAN_BENCH_DEBUG"):
        print("[bench] per-step ms:", " ".join("%.3f" % x for x in per_step), file=sys.stderr)
    total_ms = torch.tensor([sum(per_step)], dtype=torch.float64, device=wl.dev)
    wl.dp_timeline = None
    if wl.distributed:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        # Where the step of rank r spends its time around the one collective: [own step, stream time inside the
        # all-reduce].  A rank that arrives early waits inside the collective for the last one, so the SMALLEST
        # all-reduce time over ranks is the cost of the collective itself and the spread is rank skew.
        evs = wl.trainer.collective_events
        wl.trainer.collective_events = None
        mine = torch.tensor([sum(per_step) / steps, sum(a.elapsed_time(b) for a, b in evs) / max(1, len(evs))],
                            dtype=torch.float64, device=wl.dev)
        allr = [torch.zeros_like(mine) for _ in range(wl.world)]
        dist.all_gather(allr, mine)
        rows = [[round(float(x), 4) for x in t.tolist()] for t in allr]
        wl.dp_timeline = {"per_rank_ms": {"step": [r[0] for r in rows], "in_allreduce": [r[1] for r in rows]},
                          "collective_ms": min(r[1] for r in rows), "max_wait_in_collective_ms": max(r[1] for r in rows),
                          "compute_ms_slowest_rank": max(r[0] - r[1] for r in rows),
                          "compute_ms_fastest_rank": min(r[0] - r[1] for r in rows),
                          "note": "one NCCL all-reduce of the flat buffer (gradients + loss normalisers) per step; "
                                  "collective_ms = smallest stream time inside it over the ranks (the rank that arrives last "
                                  "waits for nobody); step(N) - step(1) = collective + skew of the slowest rank's compute"}
    return total_ms.item() / steps, int(launches)


def measure_e2e(wl, steps, flush, barrier, dist):
    """A training loop as one would write it: every step copies its batch from pinned host memory (H2D, async) and
    copies its loss to pinned host memory (D2H, async); the host reads the loss of step i-1 after it has enqueued
    step i, so that the device never waits for the enqueueing thread.  All copies of all steps are inside the
    timed region, which ends with a full synchronize after the last loss has been read."""
    loss_pin = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ev = [torch.cuda.Event() for _ in range(2)]
    state = {"loss": float("nan")}

    def loop(n):
        for i in range(n):
            loss = wl.step_from_host()
            loss_pin[i % 2].copy_(loss, non_blocking=True)
            loss_ev[i % 2].record()
            if i > 0:
                loss_ev[(i - 1) % 2].synchronize()
                state["loss"] = float(loss_pin[(i - 1) % 2])
        loss_ev[(n - 1) % 2].synchronize()
        state["loss"] = float(loss_pin[(n - 1) % 2])

    loop(2)
    flush.zero_()
    barrier()
    t0 = time.perf_counter()
    loop(steps)
    torch.cuda.synchronize()
    dt = torch.tensor([(time.perf_counter() - t0) / steps], dtype=torch.float64, device=wl.dev)
    if wl.distributed:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    h2d = sum(v.numel() * v.element_size() for v in wl.pinned.values()) + 4 * B_PER_GPU
    return dt.item(), h2d, state["loss"]


def measure_decode(wl, flush, barrier, dist, n_dec=10):
    """Greedy decoding of B=200 sequences, max_decoding_steps=120, EOS unreachable (every sequence runs all 121 steps:
    the worst case).  Every rank decodes its own batch (replicas: no collective on the path)."""
    model, dev, world = wl.model, wl.dev, wl.world
    model.eval()
    r = wl.resident
    for _ in range(3):
        model.greedy_decode(r["commands"], wl.cmd_len, r["situations"], 120, 1, -1)
    barrier()
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(n_dec)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(n_dec)]
    for i in range(n_dec):
        flush.zero_()
        ev0[i].record()
        model.greedy_decode(r["commands"], wl.cmd_len, r["situations"], 120, 1, -1)
        ev1[i].record()
    barrier()
    ms = torch.tensor([sum(a.elapsed_time(b) for a, b in zip(ev0, ev1)) / n_dec], dtype=torch.float64, device=dev)
    # e2e: pinned host inputs -> public API -> tokens and lengths back in pinned host memory, wall clock
    p = wl.pinned
    tok_pin = torch.empty(B_PER_GPU, 121, dtype=torch.int64).pin_memory()
    len_pin = torch.empty(B_PER_GPU, dtype=torch.int32).pin_memory()

    def e2e_once():
        c = p["commands"].to(dev, non_blocking=True)
        s = p["situations"].to(dev, non_blocking=True)
        out = model.greedy_decode(c, wl.cmd_len, s, 120, 1, -1)
        tok_pin.copy_(out["tokens"], non_blocking=True)
        len_pin.copy_(out["lengths"], non_blocking=True)
        torch.cuda.current_stream().synchronize()      # the caller reads the tokens now
        return int(len_pin[0])

    e2e_once()
    flush.zero_()
    barrier()
    t0 = time.perf_counter()
    for _ in range(n_dec):
        e2e_once()
    e2e_s = torch.tensor([(time.perf_counter() - t0) / n_dec], dtype=torch.float64, device=dev)
    if wl.distributed:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    model.train()
    dec_ms, n_seq = ms.item(), world * B_PER_GPU
    flop = B_PER_GPU * (DECODE_FLOP_FIXED + 121 * DECODE_FLOP_PER_STEP)
    return {"seqs_per_sec": n_seq / (dec_ms * 1e-3), "steps_per_sec": n_seq * 121 / (dec_ms * 1e-3),
            "ms_per_batch": dec_ms, "batch": B_PER_GPU, "steps_per_seq": 121, "n_gpus": world,
            "scaling": "replicas only: every rank decodes its own batch of 200, no collective (time = max over ranks)",
            "us_per_step": 1e3 * dec_ms / 121,
            "achieved_tflops_per_gpu": flop / (dec_ms * 1e-3) / 1e12,
            "e2e": {"value": n_seq / e2e_s.item(), "unit": "sequences/s", "ms_per_batch": 1e3 * e2e_s.item(),
                    "h2d_bytes_per_batch": p["commands"].numel() * 8 + p["situations"].numel() * 4 + 4 * B_PER_GPU,
                    "d2h_bytes_per_batch": tok_pin.numel() * 8 + len_pin.numel() * 4,
                    "note": "wall clock per call: commands + situations from pinned host memory, greedy_decode "
                            "through Model, tokens [200,121] int64 and lengths copied back to pinned host memory, "
                            "stream synchronize"},
            "note": "EOS unreachable: every sequence runs all 121 steps (worst case); one kernel launch covers all steps "
                    "(same cluster sweep as training, token feedback inside the kernel)"}


def main():
    json_out = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-gpu"])
    ap.add_argument("--workload", default="comp", choices=sorted(WORKLOADS),
                    help="comp = BASELINE.json configs[1] (the headline); comp_aux / tlen = configs[2] / configs[4]")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-decode", action="store_true")
    ap.add_argument("--no-gpu-eager", action="store_true")
    ap.add_argument("--with-decode", action="store_true", help="reference arm: also time predict() at batch size 1")
    ap.add_argument("--decode-seqs", type=int, default=16)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args, json_out)
        return
    if args.impl == "reference-gpu":
        run_reference_gpu_eager(args, json_out)
        return
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    import multimodal_seq2seq_gscan_b200 as pkg

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed:
        dist.init_process_group("nccl", device_id=dev)
    lib = pkg.load()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    wl = Workload(args.workload, dev, rank, world, distributed)

    # ---- value: device-resident, per-step CUDA events, L2 flush between steps -------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_per_step, launches = measure_training(wl, args.steps, args.warmup, flush, barrier, dist, lib)
    value = world * B_PER_GPU / (ms_per_step * 1e-3)

    # ---- e2e: pinned host buffers -> public API -> loss read back ----------------------------------
    e2e_steps = max(args.steps, 100)
    e2e_s, h2d, loss_host = measure_e2e(wl, e2e_steps, flush, barrier, dist)
    e2e_value = world * B_PER_GPU / e2e_s
    clocks = sampler.stop() if rank == 0 else None

    # ---- roofline: library-recorded CUDA events around each stage, same stream ----------------------
    lib.gscan_profile(1)
    stage_ms = np.zeros(9)
    n_prof = max(3, min(args.steps, 10))
    buf = (torch.zeros(9, dtype=torch.float32)).numpy()
    for _ in range(n_prof):
        flush.zero_()
        wl.step_resident()
        torch.cuda.synchronize()
        lib.gscan_profile_read(buf.ctypes.data)
        stage_ms += buf
    lib.gscan_profile(0)
    stage_ms /= n_prof
    stage_names = ["encoder_side", "dec_prelude", "dec_fwd_sweep", "out_proj", "", "out_proj_bwd", "dec_bwd_sweep",
                   "dec_wgrad_gemms", "encoder_side_bwd"]
    stages = {n: round(float(ms), 4) for n, ms in zip(stage_names, stage_ms) if n}

    # ---- greedy decode (BASELINE.json configs[3]) -------------------------------------------------------
    decode = None if args.no_decode else measure_decode(wl, flush, barrier, dist)

    # ---- N > 1: the other data-parallel shape BASELINE.json names (configs[4], target_length_split) -----------
    extra = None
    if distributed and args.workload != "tlen":
        wl2 = Workload("tlen", dev, rank, world, distributed)
        ms2, _ = measure_training(wl2, min(args.steps, 10), 3, flush, barrier, dist, lib)
        extra = {"tlen": {"workload": WORKLOADS["tlen"], "value": world * B_PER_GPU / (ms2 * 1e-3), "unit": "examples/s",
                          "ms_per_step": ms2, "steps": min(args.steps, 10), "n_gpus": world}}
        del wl2

    if rank == 0:
        peaks = read_peaks()
        Tt = wl.host["targets"].shape[1]
        bwd_ms, fwd_ms = stages["dec_bwd_sweep"], stages["dec_fwd_sweep"]
        achieved_tflops = B_PER_GPU * Tt * SWEEP_BWD_FLOP / (bwd_ms * 1e-3) / 1e12
        sm_mhz = (clocks or {}).get("sm_mhz") or peaks["sm_max_mhz"]
        # legacy tensor path (mma.sync m16n8k8 tf32): 512 MAC/clk/SM measured (tools/ubench_mma.cu); the sweeps
        # spend 3 MMAs per fp32-accurate product (3xTF32), so the fp32-equivalent ceiling is a third of that
        tf32_mma_peak = 148 * 512 * 2 * sm_mhz * 1e6 / 1e12
        ncu = read_ncu_profile()
        model = ncu.get("latency_model") or {}
        bwd_us, fwd_us = 1e3 * bwd_ms / Tt, 1e3 * fwd_ms / Tt
        roofline = {
            "kernel": "v3::dec_bwd_v3_kernel (BPTT cluster sweep over all 121 steps, one launch)", "bound": "latency",
            "achieved": achieved_tflops, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
            "frac": achieved_tflops / peaks["bf16_tflops"], "peak_source": peaks["source"],
            "traffic": ncu.get("bwd_dram_bytes"),
            "note": "recurrence: 121 dependent steps x ~12 dependent phases per step inside a 5-CTA cluster; neither the "
                    "tensor pipe nor HBM bounds it (see `ncu`), the dependent chain does.  Figure of merit: "
                    "us_per_decoder_step against critical_path_us_model (per-phase latency model, "
                    "profiles/r02_ncu_sweeps.md).  Mat-vecs run on mma.sync tf32 in split precision (3 MMAs per "
                    "fp32-accurate product).",
            "kernel_ms": bwd_ms, "us_per_decoder_step": bwd_us,
            "critical_path_us_model": model.get("bwd_us"),
            "frac_of_model": (model["bwd_us"] / bwd_us) if model.get("bwd_us") else None,
            "mma_sync_tf32_peak_tflops": tf32_mma_peak, "frac_mma_sync_tf32_3x": 3 * achieved_tflops / tf32_mma_peak,
            "fwd_kernel": "v3::dec_fwd_v3_kernel", "fwd_sweep_ms": fwd_ms, "fwd_us_per_decoder_step": fwd_us,
            "fwd_critical_path_us_model": model.get("fwd_us"),
            "fwd_frac_of_model": (model["fwd_us"] / fwd_us) if model.get("fwd_us") else None,
            "fwd_achieved_tflops": B_PER_GPU * Tt * SWEEP_FWD_FLOP / (fwd_ms * 1e-3) / 1e12,
            "fwd_traffic": ncu.get("fwd_dram_bytes"),
            "ncu": ncu.get("summary"), "ncu_file": ncu.get("file"),
            "stage_ms": stages,
            "whole_step_tflops": value * STEP_FLOP_PER_EXAMPLE / 1e12,
            "whole_step_frac_of_peak": value * STEP_FLOP_PER_EXAMPLE / 1e12 / peaks["bf16_tflops"],
        }
        if decode is not None:
            decode["frac_of_peak"] = decode["achieved_tflops_per_gpu"] / peaks["bf16_tflops"]
            decode["critical_path_us_model"] = model.get("fwd_us")
            decode["frac_of_model"] = (model["fwd_us"] / decode["us_per_step"]) if model.get("fwd_us") else None
        cpu_baseline = reference_gpu = None
        if world == 1 and not args.no_cpu_baseline:
            ref = _side_leg(["--impl", "reference", "--steps", "7", "--warmup", "2", "--workload", args.workload] +
                            ([] if args.no_decode else ["--with-decode"]), 600, hide_cuda=True)
            cpu_baseline = ref.get("cpu_baseline") or {"error": ref.get("error")}
            if decode is not None and "decode" in ref:
                d = ref["decode"]
                decode["cpu_baseline"] = {"value": d["seqs_per_sec"], "unit": "sequences/s", "cores": d["cores"],
                                          "kind": d["kind"], "sample": d["sample"]}
        if world == 1 and not args.no_gpu_eager:
            reference_gpu = _side_leg(["--impl", "reference-gpu", "--steps", "4", "--warmup", "2", "--workload",
                                       args.workload], 600, hide_cuda=False)
        line = {
            "metric": "train_examples_per_sec", "value": value, "unit": "examples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(args.workload, world),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "examples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": 1e3 * e2e_s, "last_loss": loss_host, "steps": e2e_steps,
                    "note": "wall clock over consecutive steps; per step: 3 H2D copies from pinned memory, train_step "
                            "through Model/FusedTrainer, async D2H of the loss into pinned memory, read by the host one "
                            "step later (no per-step device drain); L2 is flushed once "
                            "before the loop and each step's ~300 MB workspace exceeds L2"},
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "reference_gpu_eager": reference_gpu,
            "decode": decode,
            "extra_workloads": extra,
            "dp_timeline": getattr(wl, "dp_timeline", None),
        }
        print(json.dumps(line), file=json_out, flush=True)
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
