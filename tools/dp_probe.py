"""2-rank probe: where does the data-parallel step spend its extra time? (torchrun --nproc-per-node 2 tools/dp_probe.py)"""
import os, sys, time, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, multimodal_seq2seq_gscan_b200 as pkg
from multimodal_seq2seq_gscan_b200 import synthetic as O, dp
from multimodal_seq2seq_gscan_b200.trainer import FusedTrainer
rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev); pkg.load()
cfg = bench.bench_cfg()
model = pkg.Model(**O.model_kwargs(cfg)).to(dev)
model.load_state_dict(O.full_state_dict(O.synthetic_params(cfg, 1234)), strict=True)
tr = FusedTrainer(model, distributed=True)
host = bench.make_host_batch(cfg, 1235 + rank)
res = {k: torch.from_numpy(np.ascontiguousarray(host[k])).to(dev) for k in ("commands", "situations", "targets")}
def step(): return tr.train_step(res["commands"], host["cmd_lengths"], res["situations"], res["targets"], host["tgt_lengths"])
def timeit(fn, n=20):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
def hosttime(fn, n=20):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    torch.cuda.synchronize()
    return 1e3 * float(np.median(ts))
t_full = timeit(step)
h_full = hosttime(step)
orig = dp.allreduce_flat_gradient
dp.allreduce_flat_gradient = lambda g, group=None: None
t_noar = timeit(step)
dp.allreduce_flat_gradient = orig
orig2 = dp.start_count_allreduce
class W:
    def wait(self): pass
dp.start_count_allreduce = lambda t, p, g=None: (dp.local_counts(t, p) * 2, W())
t_nocount = timeit(step)
dp.allreduce_flat_gradient = lambda g, group=None: None
t_none = timeit(step)
lib = pkg.load()
lib.gscan_profile(1)
buf = np.zeros(9, dtype=np.float32); acc = np.zeros(9)
for _ in range(10):
    step(); torch.cuda.synchronize(); lib.gscan_profile_read(buf.ctypes.data); acc += buf
lib.gscan_profile(0)
names = ["encoder_side", "dec_prelude", "dec_fwd_sweep", "out_proj", "", "out_proj_bwd", "dec_bwd_sweep", "dec_wgrad_gemms", "encoder_side_bwd"]
if rank == 0: print("stages(no collectives):", {n: round(float(v) / 10, 3) for n, v in zip(names, acc) if n})
tr2 = FusedTrainer(model, distributed=False)
def step2(): return tr2.train_step(res["commands"], host["cmd_lengths"], res["situations"], res["targets"], host["tgt_lengths"])
t_plain = timeit(step2)
if rank == 0: print(f"non-distributed trainer in the same process: {t_plain:.3f} ms")
gl = dp.global_loss
dp.allreduce_flat_gradient = lambda g, group=None: None
dp.start_count_allreduce = lambda t, p, g=None: (None, W())
dp.global_loss = lambda nll, n_tok, aux, B, w, counts: nll
v1 = timeit(step)
dp.start_count_allreduce = lambda t, p, g=None: (dp.local_counts(t, p), W())
v2 = timeit(step)
Bc = torch.tensor(200.0, device=dev)
def lc_b(t, p): return torch.stack(((t[:, 1:] != p).sum(dtype=torch.float32), Bc))
dp.start_count_allreduce = lambda t, p, g=None: (lc_b(t, p), W())
v2b = timeit(step)
def lc_c(t, p):
    n = (t[:, 1:] != p).sum()
    return None
dp.start_count_allreduce = lambda t, p, g=None: (lc_c(t, p), W())
v2c = timeit(step)
def lc_d(t, p):
    c = torch.empty(2, dtype=torch.float32, device=t.device); c[1] = 200.0
    return c
dp.start_count_allreduce = lambda t, p, g=None: (lc_d(t, p), W())
v2d = timeit(step)
if rank == 0: print(f"V2b stack(sum,const) {v2b:.3f} | V2c only ne+sum {v2c:.3f} | V2d only empty+setitem float {v2d:.3f}")
cst = torch.tensor(0.5, device=dev)
dp.start_count_allreduce = lambda t, p, g=None: (None, W())
dp.global_loss = lambda nll, n_tok, aux, B, w, counts: nll * cst
v3 = timeit(step)
dp.global_loss = lambda nll, n_tok, aux, B, w, counts: nll * (n_tok.detach() / 7.0)
v4 = timeit(step)
if rank == 0: print(f"V1 plain loss {v1:.3f} | V2 +local_counts {v2:.3f} | V3 loss*const {v3:.3f} | V4 loss*(n_tok/7) {v4:.3f}")
dp.global_loss = gl; dp.allreduce_flat_gradient = orig; dp.start_count_allreduce = orig2
g = torch.zeros(440320, device=dev)
t_ar = timeit(lambda: dist.all_reduce(g))
if rank == 0:
    print(f"host enqueue per step {h_full:.3f} ms, OMP_NUM_THREADS={os.environ.get('OMP_NUM_THREADS')}, torch threads {torch.get_num_threads()}, cpus {len(os.sched_getaffinity(0))}")
    print(f"full {t_full:.3f} ms | no grad allreduce {t_noar:.3f} | no count allreduce {t_nocount:.3f} | neither {t_none:.3f} | bare allreduce 1.76MB back-to-back {t_ar*1e3:.1f} us")
dist.destroy_process_group()
