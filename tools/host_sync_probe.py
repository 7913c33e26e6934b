"""Does the host run ahead of the GPU in the training loop?  Host time per train_step call without any explicit
synchronisation, and PyTorch's sync-debug warnings for implicit ones.  usage: python tools/host_sync_probe.py"""
import os, sys, time, warnings
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
wl = bench.Workload("comp", dev, 0, 1, False)
for _ in range(10):
    wl.step_resident()
torch.cuda.synchronize()
torch.cuda.set_sync_debug_mode("warn")
with warnings.catch_warnings(record=True) as w:
    warnings.simplefilter("always")
    wl.step_resident()
    for x in w:
        print("SYNC WARNING:", str(x.message)[:200], x.filename, x.lineno)
torch.cuda.set_sync_debug_mode("default")
torch.cuda.synchronize()
ts = []
t_all = time.perf_counter()
for _ in range(40):
    t0 = time.perf_counter()
    wl.step_resident()
    ts.append((time.perf_counter() - t0) * 1e3)
t_host = (time.perf_counter() - t_all) * 1e3
torch.cuda.synchronize()
t_total = (time.perf_counter() - t_all) * 1e3
print("host ms per call:", " ".join("%.2f" % t for t in ts))
print("host loop %.1f ms for 40 steps, until GPU idle %.1f ms (%.3f ms/step)" % (t_host, t_total, t_total / 40))
