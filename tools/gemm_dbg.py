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
