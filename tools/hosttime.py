import time, torch, numpy as np, sys
sys.path.insert(0, '.')
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
for _ in range(5): step()
torch.cuda.synchronize()
ts=[]
for _ in range(20):
    torch.cuda.synchronize()
    t0=time.perf_counter(); step(); t1=time.perf_counter(); torch.cuda.synchronize(); t2=time.perf_counter()
    ts.append((t1-t0, t2-t0))
print('host enqueue ms (median):', 1e3*np.median([a for a,_ in ts]), ' total ms:', 1e3*np.median([b for _,b in ts]))
import cProfile, pstats
pr=cProfile.Profile(); pr.enable()
for _ in range(20): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
