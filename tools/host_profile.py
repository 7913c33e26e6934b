"""Where the host time of one train_step goes (cProfile over 60 steps, GPU left to run behind).
usage: python tools/host_profile.py"""
import cProfile, os, pstats, sys, io
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
wl = bench.Workload("comp", dev, 0, 1, False)
for _ in range(10):
    wl.step_resident()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(60):
    wl.step_resident()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
