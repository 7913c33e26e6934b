"""the situation CNN as a zero-skipping gather (csrc/cnn.cuh) against a tensor-core GEMM of the im2col
shape, at several input densities.  The gather's cost grows with the number of non-zeros; the GEMM's does not.
    M = B*36 = 7200 output cells, N = 150 output channels, K = (1 + 25 + 49) * 16 = 1200 taps x channels
The GEMM leg times the library's tcgen05 3xTF32 kernel on that shape (gscan_sgemm_path, path 1) and adds the time of one
memory-bound pass that would write the [7200 x 1200] fp32 im2col matrix (34.6 MB at the measured HBM rate): the two
dominant costs of a dense implicit-GEMM formulation, without its epilogue.  Prints one JSON line."""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multimodal_seq2seq_gscan_b200 as pkg
from multimodal_seq2seq_gscan_b200 import ops, synthetic as S, _lib

dev = torch.device("cuda:0")
lib = pkg.load()
cfg = dict(S.CONFIGS["comp"])
model = pkg.Model(**S.model_kwargs(cfg)).to(dev)
model.load_state_dict(S.full_state_dict(S.synthetic_params(cfg, 1)), strict=True)
B, G, C = 200, 6, 16
rng = np.random.default_rng(0)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for a, b in ev:
        flush.zero_()
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in ev])) * 1e3   # us


out = {"shape": "B=200 G=6 C=16 F=50 k3=7 (M=7200, N=150, K=1200)", "gather_us": {}}
batch = S.synthetic_batch(cfg, batch_size=B, seed=2)
sit = torch.tensor(batch["situations"], device=dev)
out["gather_us"][f"gSCAN-like ({float((sit != 0).float().mean()) * 100:.1f} % non-zero)"] = timeit(
    lambda: ops.cnn_forward(model._cfg(G), model._param_list(), sit))
for dens in (0.05, 0.25, 1.0):
    x = torch.tensor((rng.random((B, G, G, C)) < dens).astype(np.float32) * rng.normal(size=(B, G, G, C)).astype(np.float32),
                     device=dev)
    out["gather_us"][f"{int(dens * 100)} %"] = timeit(lambda: ops.cnn_forward(model._cfg(G), model._param_list(), x))
M, N, K = 7200, 150, 1200
A = torch.randn(M, K, device=dev)
W = torch.randn(N, K, device=dev)
Cm = torch.empty(M, N, device=dev)
st = torch.cuda.current_stream().cuda_stream


def gemm(path):
    rc = lib.gscan_sgemm_path(A.data_ptr(), K, 1, W.data_ptr(), 1, K, Cm.data_ptr(), N, M, N, K, None, 0, 0, 1, path, st)
    assert rc == 0, rc


out["gemm_tcgen05_3xtf32_us"] = timeit(lambda: gemm(1))
out["gemm_mma_sync_3xtf32_us"] = timeit(lambda: gemm(0))
out["im2col_write_us_at_6.5TBs"] = M * K * 4 / 6.55e12 * 1e6
ref = (A.double() @ W.double().T)
out["gemm_tcgen05_rel_err"] = float(((Cm.double() - ref).norm() / ref.norm()))
print(json.dumps(out))
