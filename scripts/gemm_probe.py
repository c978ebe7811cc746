import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g; g.build()
from zeggs_b200 import _lib, ops
dev = torch.device("cuda:0"); ops.ensure_scratch(dev)
def run(mode, M, N, K):
    gen = torch.Generator().manual_seed(1)
    if mode == 0: A, B = torch.randn(M, K, generator=gen), torch.randn(N, K, generator=gen); ref = A.double() @ B.double().T
    elif mode == 1: A, B = torch.randn(K, M, generator=gen), torch.randn(K, N, generator=gen); ref = A.double().T @ B.double()
    else: A, B = torch.randn(M, K, generator=gen), torch.randn(K, N, generator=gen); ref = A.double() @ B.double()
    Ad, Bd = A.to(dev), B.to(dev); out = torch.full((M, N), float('nan'), device=dev)
    _lib.check(_lib.lib().zeggs_gemm_f32_ctx(ops.ctx_ptr(dev), mode, M, N, K, Ad.data_ptr(), Ad.stride(0), Bd.data_ptr(), Bd.stride(0), None, out.data_ptr(), N, 0, 0, _lib.stream_ptr()), "g")
    err = (out.cpu().double() - ref).abs().max().item(); print(f"mode{mode} {M}x{N}x{K}: err {err:.3e} / max {ref.abs().max().item():.3e}  nan={torch.isnan(out).sum().item()}")
for shp in [(1,512,3402,32),(1,512,3402,99),(1,512,3402,260),(1,512,3402,264),(1,128,1536,32),(1,128,1536,260),(2,32,1536,128),(2,260,1536,128),(2,99,1536,128),
            (0,32,512,3402),(0,260,512,3402),(0,260,128,1536),(1,384,128,260),(2,260,128,384),(1,128,384,260),(2,260,384,128)]:
    run(*shp)
