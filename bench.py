#!/usr/bin/env python
"""Benchmark of the gSCAN seq2seq hot path on B200 (BASELINE.json metric: train examples/sec at
batch 200 per GPU, plus batched greedy-decode sequences/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full training iteration of the compositional_splits shape (BASELINE.json configs[1]:
grid 6, B=200 per GPU, k=7, E=25, H=100, Ti=10, Tt=121 - all padded steps computed as the reference
does, paper dropout 0.3/0.3/0.1, no auxiliary task): forward + NLL loss + backward + (gradient
all-reduce when N>1) + Adam.  Rank 0 prints ONE JSON line.

  value     examples/s with the batch already resident in HBM; every step is bracketed by its own
            CUDA events on the launching stream and L2 is flushed (256 MiB memset) between steps;
            the per-rank totals are reduced with MAX over ranks.
  e2e       the same step driven from pinned HOST buffers through the public API
            (Model/FusedTrainer): per step the H2D copies of the batch and an async D2H copy of the
            loss, which the host reads one step later; wall clock over consecutive steps with a
            device synchronize at both ends.
  roofline  the dominant kernel (decoder backward cluster sweep) timed live with CUDA events recorded by
            the library on the same stream (gscan_profile); the sweep is a latency / synchronisation bound
            recurrence (121 dependent steps), so its fraction of the measured bf16 tensor peak is tiny by
            construction - the object also carries microseconds per decoder step, the fraction of the
            mma.sync tf32 ceiling, and the DRAM traffic per launch from the committed ncu capture.
  cpu_baseline / --impl reference
            the CPU port of the reference (oracle/gscan_oracle.py: same per-step PyTorch-eager
            structure) on all host cores, same shape, same step definition.  /root/reference itself
            is Python source that does not exist on the GPU box.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
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
WORKLOAD = WORKLOADS["comp"]
_WORKLOAD_KEY = "comp"
SEED = 1234
B_PER_GPU = 200
# algorithmic FLOPs (2 per MAC) per example per decoder step inside the recurrent sweeps:
# forward: the four dependent mat-vec stages (600+500+100+400 rows x 100) + both attentions (46 keys x 100 x 2)
SWEEP_FWD_FLOP = 2 * (1600 * 100 + 46 * 100 * 2)
# backward: the transposed mat-vecs (same 160k MAC) + attention backward (46 keys x 100 x 4)
SWEEP_BWD_FLOP = 2 * (1600 * 100 + 46 * 100 * 4)
# whole training step per example: 3 x (7.62M + 121 x 0.506M)
STEP_FLOP_PER_EXAMPLE = 206.5e6


def bench_cfg():
    from multimodal_seq2seq_gscan_b200 import synthetic
    cfg = dict(synthetic.CONFIGS["tlen" if _WORKLOAD_KEY == "tlen" else "comp"])
    cfg.update(encoder_dropout_p=0.3, decoder_dropout_p=0.3, cnn_dropout_p=0.1,
               auxiliary_task=(_WORKLOAD_KEY == "comp_aux"))
    return cfg


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
            # the sweep is timed inside a long step: the sustained figure is the one that applies, when the file has it
            for cand in (key + "_sustained", key):
                if cand in p:
                    out[key] = number(p[cand])
                    break
        return out
    except Exception as exc:      # a malformed file must not take the bench line down
        return dict(fallback, source=f"fallback (MEASURED_PEAKS.json unreadable: {type(exc).__name__})")


def read_ncu_profile():
    """DRAM bytes per launch of the two sweeps from the committed `ncu --set full` capture (profiles/)."""
    path = os.path.join(ROOT, "profiles", "r01_v3_ncu_sweeps.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
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


def make_host_batch(cfg, seed):
    from multimodal_seq2seq_gscan_b200 import synthetic
    return synthetic.synthetic_batch(cfg, batch_size=B_PER_GPU, seed=seed,
                                     min_tgt_len=17 if _WORKLOAD_KEY == "tlen" else 3)


# --------------------------------------------------------------------------------------------
# CPU reference arm (oracle port) - the only place bench.py executes oracle/
# --------------------------------------------------------------------------------------------
def cpu_reference_steps(cfg, steps, warmup, threads=None):
    """forward + loss + backward + Adam on the host cores with the oracle port (train mode: dropout
    masks drawn with torch's CPU RNG).  Returns (examples/s from the median step, list of seconds)."""
    from oracle import gscan_oracle as O
    threads = threads or len(os.sched_getaffinity(0))
    torch.set_num_threads(threads)
    params = {k: v.clone().requires_grad_(True) for k, v in O.synthetic_params(cfg, SEED).items()}
    opt = torch.optim.Adam(list(params.values()), lr=1e-3)
    batch = make_host_batch(cfg, SEED + 1)
    commands, targets = torch.tensor(batch["commands"]), torch.tensor(batch["targets"])
    situations = torch.tensor(batch["situations"])
    B, Ti, Tt = commands.shape[0], commands.shape[1], targets.shape[1]
    M, D, E, H = cfg["grid_size"] ** 2, 3 * cfg["cnn_hidden_num_channels"], cfg["embedding_dimension"], cfg["decoder_hidden_size"]

    def mask(shape, p):
        return None if p <= 0 else torch.empty(shape).bernoulli_(1 - p) / (1 - p)

    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        drop = {"cnn": mask((B, M, D), cfg["cnn_dropout_p"]), "enc": mask((B, Ti, E), cfg["encoder_dropout_p"]),
                "dec": mask((B, Tt, H), cfg["decoder_dropout_p"])}
        logp, aux = O.model_forward(params, commands, batch["cmd_lengths"], situations, targets,
                                    cfg["conditional_attention"], cfg["auxiliary_task"], dropout=drop)
        loss = O.nll_loss(logp, targets)
        if cfg["auxiliary_task"]:
            loss = loss + 0.3 * O.aux_nll_loss(aux, torch.tensor(batch["target_positions"]))
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    med = statistics.median(times)
    return B / med, times, threads


def run_reference_arm(args, json_out):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = bench_cfg()
    steps = max(1, min(args.steps, 20))
    warmup = max(1, min(args.warmup, 3))
    value, times, threads = cpu_reference_steps(cfg, steps, warmup)
    sample = f"{steps} full steps (B=200, Tt=121, fwd+loss+bwd+Adam) after {warmup} warm-up, median"
    line = {
        "impl": "reference", "metric": "train_examples_per_sec", "value": value, "unit": "examples/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * statistics.median(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "device": "host CPU", "global_batch": B_PER_GPU},
        "cpu_baseline": {"value": value, "unit": "examples/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "examples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=json_out, flush=True)


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


def main():
    json_out = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="comp", choices=sorted(WORKLOADS),
                    help="comp = BASELINE.json configs[1] (the headline); comp_aux / tlen = configs[2] / configs[4]")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-decode", action="store_true")
    args = ap.parse_args()
    global WORKLOAD, _WORKLOAD_KEY
    _WORKLOAD_KEY, WORKLOAD = args.workload, WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference_arm(args, json_out)
        return
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    import multimodal_seq2seq_gscan_b200 as pkg
    from multimodal_seq2seq_gscan_b200 import synthetic as O
    from multimodal_seq2seq_gscan_b200.trainer import FusedTrainer

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

    cfg = bench_cfg()
    model = pkg.Model(**O.model_kwargs(cfg)).to(dev)
    model.load_state_dict(O.full_state_dict(O.synthetic_params(cfg, SEED)), strict=True)
    trainer = FusedTrainer(model, distributed=distributed)
    host = make_host_batch(cfg, SEED + 1 + rank)
    pinned = {k: torch.from_numpy(np.ascontiguousarray(host[k])).pin_memory()
              for k in ("commands", "situations", "targets")}
    resident = {k: v.to(dev) for k, v in pinned.items()}
    cmd_len, tgt_len = host["cmd_lengths"], host["tgt_lengths"]
    positions = torch.from_numpy(host["target_positions"]).to(dev) if cfg["auxiliary_task"] else None
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step_resident():
        return trainer.train_step(resident["commands"], cmd_len, resident["situations"], resident["targets"], tgt_len,
                                  positions)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_resident()
    barrier()

    # ---- value: device-resident, per-step CUDA events, L2 flush between steps -------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    launches0 = lib.gscan_launch_count()
    barrier()
    for i in range(args.steps):
        flush.zero_()
        starts[i].record()
        step_resident()
        ends[i].record()
    barrier()
    launches = lib.gscan_launch_count() - launches0
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    ms_per_step = total_ms.item() / args.steps
    value = world * B_PER_GPU / (ms_per_step * 1e-3)

    # ---- e2e: pinned host buffers -> public API -> loss read back ----------------------------------
    # A training loop as one would write it: every step copies its batch from pinned host memory (H2D, async) and
    # copies its loss to pinned host memory (D2H, async); the host reads the loss of step i-1 after it has enqueued
    # step i, so that the device never waits for the enqueueing thread.  All copies of all steps are inside the
    # timed region, which ends with a full synchronize after the last loss has been read.
    e2e_steps = max(3, min(args.steps, 20))
    loss_pin = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ev = [torch.cuda.Event() for _ in range(2)]
    loss_host = float("nan")

    def e2e_loop(n):
        nonlocal loss_host
        for i in range(n):
            c = pinned["commands"].to(dev, non_blocking=True)
            s = pinned["situations"].to(dev, non_blocking=True)
            t = pinned["targets"].to(dev, non_blocking=True)
            loss = trainer.train_step(c, cmd_len, s, t, tgt_len, positions)
            loss_pin[i % 2].copy_(loss, non_blocking=True)
            loss_ev[i % 2].record()
            if i > 0:
                loss_ev[(i - 1) % 2].synchronize()
                loss_host = float(loss_pin[(i - 1) % 2])
        loss_ev[(n - 1) % 2].synchronize()
        loss_host = float(loss_pin[(n - 1) % 2])

    e2e_loop(2)
    flush.zero_()
    barrier()
    t0 = time.perf_counter()
    e2e_loop(e2e_steps)
    torch.cuda.synchronize()
    e2e_dt = (time.perf_counter() - t0) / e2e_steps
    e2e_t = torch.tensor([e2e_dt], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = world * B_PER_GPU / e2e_t.item()
    h2d = sum(v.numel() * v.element_size() for v in pinned.values()) + 4 * B_PER_GPU
    clocks = sampler.stop() if rank == 0 else None

    # ---- roofline: library-recorded CUDA events around each stage, same stream ----------------------
    lib.gscan_profile(1)
    stage_ms = np.zeros(9)
    n_prof = max(3, min(args.steps, 10))
    buf = (torch.zeros(9, dtype=torch.float32)).numpy()
    for _ in range(n_prof):
        flush.zero_()
        step_resident()
        torch.cuda.synchronize()
        lib.gscan_profile_read(buf.ctypes.data)
        stage_ms += buf
    lib.gscan_profile(0)
    stage_ms /= n_prof
    stage_names = ["encoder_side", "dec_prelude", "dec_fwd_sweep", "out_proj", "", "out_proj_bwd", "dec_bwd_sweep",
                   "dec_wgrad_gemms", "encoder_side_bwd"]
    stages = {n: round(float(ms), 4) for n, ms in zip(stage_names, stage_ms) if n}

    # ---- greedy decode (BASELINE.json configs[3]): B=200, max_decoding_steps=120, EOS unreachable ---
    decode = None
    if not args.no_decode:
        model.eval()
        for _ in range(3):
            model.greedy_decode(resident["commands"], cmd_len, resident["situations"], 120, 1, -1)
        torch.cuda.synchronize()
        n_dec = 10
        ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(n_dec)]
        ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(n_dec)]
        for i in range(n_dec):
            flush.zero_()
            ev0[i].record()
            out = model.greedy_decode(resident["commands"], cmd_len, resident["situations"], 120, 1, -1)
            ev1[i].record()
        torch.cuda.synchronize()
        dec_ms = sum(a.elapsed_time(b) for a, b in zip(ev0, ev1)) / n_dec
        decode = {"seqs_per_sec": B_PER_GPU / (dec_ms * 1e-3), "steps_per_sec": B_PER_GPU * 121 / (dec_ms * 1e-3),
                  "ms_per_batch": dec_ms, "batch": B_PER_GPU, "steps_per_seq": 121, "n_gpus": 1,
                  "note": "EOS unreachable: every sequence runs all 121 steps (worst case)"}
        model.train()

    if rank == 0:
        peaks = read_peaks()
        Tt = host["targets"].shape[1]
        bwd_ms = stages["dec_bwd_sweep"]
        fwd_ms = stages["dec_fwd_sweep"]
        achieved_tflops = B_PER_GPU * Tt * SWEEP_BWD_FLOP / (bwd_ms * 1e-3) / 1e12
        sm_mhz = (clocks or {}).get("sm_mhz") or peaks["sm_max_mhz"]
        # legacy tensor path (mma.sync m16n8k8 tf32): 512 MAC/clk/SM measured (tools/ubench_mma.cu); the sweeps
        # spend 3 MMAs per fp32-accurate product (3xTF32), so the fp32-equivalent ceiling is a third of that
        tf32_mma_peak = 148 * 512 * 2 * sm_mhz * 1e6 / 1e12
        ncu = read_ncu_profile()
        roofline = {
            "kernel": "v3::dec_bwd_v3_kernel (BPTT cluster sweep over all 121 steps, one launch)", "bound": "tensor",
            "achieved": achieved_tflops, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
            "frac": achieved_tflops / peaks["bf16_tflops"], "peak_source": peaks["source"],
            "traffic": ncu.get("bwd_dram_bytes"),
            "note": "recurrence: 121 dependent steps x ~12 dependent phases; latency/synchronisation bound, not tensor or "
                    "HBM bound (ncu: tensor pipe ~7 % active, issue slots ~40 % busy, 27 % of stall samples at block "
                    "barriers).  Figure of merit: us_per_decoder_step.  Mat-vecs run on mma.sync tf32 in split "
                    "precision (3 MMAs per fp32-accurate product).",
            "mma_sync_tf32_peak_tflops": tf32_mma_peak, "frac_mma_sync_tf32_3x": 3 * achieved_tflops / tf32_mma_peak,
            "kernel_ms": bwd_ms, "us_per_decoder_step": 1e3 * bwd_ms / Tt,
            "fwd_kernel": "v3::dec_fwd_v3_kernel", "fwd_sweep_ms": fwd_ms, "fwd_us_per_decoder_step": 1e3 * fwd_ms / Tt,
            "fwd_achieved_tflops": B_PER_GPU * Tt * SWEEP_FWD_FLOP / (fwd_ms * 1e-3) / 1e12,
            "fwd_traffic": ncu.get("fwd_dram_bytes"),
            "ncu": ncu.get("summary"),
            "stage_ms": stages,
            "whole_step_tflops": value * STEP_FLOP_PER_EXAMPLE / 1e12,
        }
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            v, times, threads = cpu_reference_steps(cfg, steps=7, warmup=2)
            cpu_baseline = {"value": v, "unit": "examples/s", "cores": threads, "kind": "port",
                            "sample": "7 full steps (B=200, Tt=121, fwd+loss+bwd+Adam) after 2 warm-up; median "
                                      f"{statistics.median(times):.3f} s, min {min(times):.3f} s"}
        line = {
            "metric": "train_examples_per_sec", "value": value, "unit": "examples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": world * B_PER_GPU,
                       "parallelism": f"dp{world}" if world > 1 else "single",
                       "step": "forward + NLL + backward + " + ("gradient all-reduce (NCCL) + " if world > 1 else "")
                               + "fused Adam",
                       "l2": "flushed between timed steps (256 MiB memset); the step's own workspace (~300 MB) "
                             "also exceeds L2"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "examples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": 1e3 * e2e_t.item(), "last_loss": loss_host, "steps": e2e_steps,
                    "note": "wall clock over consecutive steps; per step: 3 H2D copies from pinned memory, train_step "
                            "through Model/FusedTrainer, async D2H of the loss into pinned memory, read by the host one "
                            "step later (no per-step device drain); L2 is flushed once "
                            "before the loop and each step's ~300 MB workspace exceeds L2"},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "decode": decode,
        }
        print(json.dumps(line), file=json_out, flush=True)
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
