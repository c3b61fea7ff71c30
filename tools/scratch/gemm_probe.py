import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from stylish_tts_b200 import _lib as L, diffusion as DF
d = torch.device("cuda:0")
for M, N, K in [(128,128,64),(128,128,128),(128,128,192),(128,128,256),(128,128,512),(128,128,1024),(256,256,64),(128,256,192),(128,128,2048), (128,128,320)]:
    g = torch.Generator().manual_seed(1)
    a = torch.randn(M, K, generator=g); w = torch.randn(N, K, generator=g) / math.sqrt(K)
    ref = a.double() @ w.double().t()
    out = torch.zeros(M, N, device=d)
    L.call("sty_gemm_split_fwd", DF._planes(a.to(d)).data_ptr(), DF._planes(w.to(d)).data_ptr(), None, None, out.data_ptr(), None, M, N, K, 0, L.stream_ptr())
    torch.cuda.synchronize()
    e = float((out.cpu().double() - ref).norm() / ref.norm())
    # which k-blocks contributed? compare against partial sums
    best = None
    for kb in range(K // 64 + 1):
        part = a[:, :kb*64].double() @ w[:, :kb*64].double().t()
        ee = float((out.cpu().double() - part).norm() / (ref.norm()))
        if best is None or ee < best[1]: best = (kb, ee)
    print(M, N, K, "err", f"{e:.3e}", "closest partial sum: first", best[0], "k-blocks", f"{best[1]:.2e}")
