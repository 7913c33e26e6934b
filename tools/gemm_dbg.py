import torch, sys
sys.path.insert(0, '/root/repo')
import multimodal_seq2seq_gscan_b200 as pkg
lib = pkg.load()
DEV = 'cuda'
for (M,N,K) in [(9,100,2000),(300,400,100),(128,64,4096),(64,64,32)]:
    g = torch.Generator().manual_seed(1)
    A = torch.randn(M, K, generator=g, dtype=torch.float64)
    Bm = torch.randn(K, N, generator=g, dtype=torch.float64)
    ref = A @ Bm
    Ad, Bd = A.float().to(DEV), Bm.float().to(DEV)
    ref32 = (Ad @ Bd).cpu().double()
    st = torch.cuda.current_stream().cuda_stream
    C = torch.zeros(M, N, device=DEV)
    rc = lib.gscan_sgemm(Ad.data_ptr(), K, 1, Bd.data_ptr(), N, 1, C.data_ptr(), N, M, N, K, None, 0, 0, st)
    err = (C.cpu().double() - ref).abs()
    print(M,N,K,'rc',rc,'max err',err.max().item(),'mean err',err.mean().item(),'ref max',ref.abs().max().item(),'torch fp32 err',(ref32-ref).abs().max().item())

# timing of the three forms at the shapes of the training step
def bench(name, M, N, K, a_rs, a_cs, b_rs, b_cs, Ashape, Bshape):
    A = torch.randn(*Ashape, device=DEV); Bm = torch.randn(*Bshape, device=DEV); C = torch.zeros(M, N, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3): lib.gscan_sgemm(A.data_ptr(), a_rs, a_cs, Bm.data_ptr(), b_rs, b_cs, C.data_ptr(), N, M, N, K, None, 0, 0, st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): lib.gscan_sgemm(A.data_ptr(), a_rs, a_cs, Bm.data_ptr(), b_rs, b_cs, C.data_ptr(), N, M, N, K, None, 0, 0, st)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1000
    print(f"{name}: M={M} N={N} K={K}: {us:.1f} us, {M*N*K/us/1e6:.1f} TMAC/s ({3*M*N*K/us/1e6/144.0*100:.0f}% of the mma.sync tf32 peak at 3 MMAs per product)")
R = 24200
bench("NT Xe", R, 400, 100, 100, 1, 1, 100, (R, 100), (400, 100))
bench("NT out", R, 100, 400, 400, 1, 1, 400, (R, 400), (100, 400))
bench("NN dU", R, 400, 100, 100, 1, 400, 1, (R, 100), (100, 400))
bench("NN post", R, 200, 400, 400, 1, 200, 1, (R, 400), (400, 200))
