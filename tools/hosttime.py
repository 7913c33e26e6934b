"""Single-GPU probe: device time per training step measured three ways, and the host enqueue time."""
import time, torch, numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, multimodal_seq2seq_gscan_b200 as pkg
from multimodal_seq2seq_gscan_b200 import synthetic as O
from multimodal_seq2seq_gscan_b200.trainer import FusedTrainer
dev = torch.device('cuda:0'); pkg.load()
cfg = bench.bench_cfg()
model = pkg.Model(**O.model_kwargs(cfg)).to(dev)
model.load_state_dict(O.full_state_dict(O.synthetic_params(cfg, 1234)), strict=True)
tr = FusedTrainer(model)
host = bench.make_host_batch(cfg, 1235)
res = {k: torch.from_numpy(np.ascontiguousarray(host[k])).to(dev) for k in ('commands','situations','targets')}
def step(): return tr.train_step(res['commands'], host['cmd_lengths'], res['situations'], res['targets'], host['tgt_lengths'])
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for _ in range(5): step()
torch.cuda.synchronize()
N = 20
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(N): step()
e1.record(); torch.cuda.synchronize()
print('back-to-back, one event pair     : %.3f ms/step' % (e0.elapsed_time(e1) / N))
s = [torch.cuda.Event(enable_timing=True) for _ in range(N)]; e = [torch.cuda.Event(enable_timing=True) for _ in range(N)]
for i in range(N):
    s[i].record(); step(); e[i].record()
torch.cuda.synchronize()
print('back-to-back, per-step events     : %.3f ms/step' % (sum(a.elapsed_time(b) for a, b in zip(s, e)) / N))
for i in range(N):
    flush.zero_(); s[i].record(); step(); e[i].record()
torch.cuda.synchronize()
print('L2 flush + per-step events (bench): %.3f ms/step' % (sum(a.elapsed_time(b) for a, b in zip(s, e)) / N))
e0.record()
for i in range(N):
    flush.zero_(); step()
e1.record(); torch.cuda.synchronize()
print('L2 flush, one event pair          : %.3f ms/step (includes the flush)' % (e0.elapsed_time(e1) / N))
ts = []
for _ in range(N):
    torch.cuda.synchronize(); t0 = time.perf_counter(); step(); ts.append(time.perf_counter() - t0)
torch.cuda.synchronize()
print('host enqueue per step             : %.3f ms' % (1e3 * float(np.median(ts))))
