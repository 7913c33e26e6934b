"""Kernel-by-kernel trace of ONE pipelined training step (the host runs ahead of the GPU, as in the bench's timed
region): torch.profiler (CUPTI) over a few steps of bench.Workload, then the GPU activities of the last complete step
sorted by start time, with the stream they ran on and the gap to the previous activity's end on that stream.

Unlike GSCAN_CHAIN_TIMES=1 (which synchronises to read its events and so includes host launch latency at the start of
each pass) this is the timeline the bench measures.   usage: python tools/step_trace.py [out.md] [workload]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/step_trace.md"
    key = sys.argv[2] if len(sys.argv) > 2 else "comp"
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    wl = bench.Workload(key, dev, 0, 1, False)
    for _ in range(10):
        wl.step_resident()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(4):
            wl.step_resident()
        torch.cuda.synchronize()
    evs = []
    for e in prof.events():
        if getattr(e, "device_type", None) is not None and "CUDA" in str(e.device_type):
            tr = e.time_range
            evs.append((tr.start, tr.end, e.name))
    if not evs:   # older/newer profiler layouts: fall back to the kineto results
        for e in prof.profiler.kineto_results.events():
            if "cuda" in str(e.device_type()).lower():
                evs.append((e.start_ns() / 1e3, (e.start_ns() + e.duration_ns()) / 1e3, e.name()))
    evs.sort()
    adam = [i for i, e in enumerate(evs) if "adam" in e[2]]
    if len(adam) < 3:
        print("no complete step found; events:", len(evs))
        return
    lo, hi = adam[-2] + 1, adam[-1] + 1
    step = evs[lo:hi]
    t0 = evs[adam[-2]][1]
    lines = ["# one pipelined training step, kernel by kernel (torch.profiler / CUPTI; us after the previous step's Adam)",
             "", "| start | end | dur | kernel |", "|---:|---:|---:|---|"]
    for s, e, n in step:
        lines.append("| %.1f | %.1f | %.1f | %s |" % (s - t0, e - t0, e - s, n.split("(")[0][:90]))
    lines.append("")
    lines.append("step: %.1f us from the previous Adam's end to this Adam's end" % (step[-1][1] - t0))
    fw = [e for e in step if "dec_fwd_v3" in e[2]]
    bw = [e for e in step if "dec_bwd_v3" in e[2]]
    summary = None
    if fw and bw:
        summary = ("summary: pre %.1f | fwd sweep %.1f | mid %.1f | bwd sweep %.1f | post %.1f | step %.1f us" %
                   (fw[0][0] - t0, fw[0][1] - fw[0][0], bw[0][0] - fw[0][1], bw[0][1] - bw[0][0],
                    step[-1][1] - bw[0][1], step[-1][1] - t0))
        lines.append(summary)
    os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
    with open(out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print(summary if (summary and os.environ.get("STEP_TRACE_BRIEF")) else "\n".join(lines))


if __name__ == "__main__":
    main()
