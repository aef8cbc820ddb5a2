"""Tiny driver for ncu: the tcgen05 scorer at the benchmark shape."""
import sys
import torch
sys.path.insert(0, ".")
from transkun_b200.LayersTransformer import sip_score
NT, T, D = 88, int(sys.argv[1]) if len(sys.argv) > 1 else 2048, 256
g = torch.Generator().manual_seed(0)
q, k, d = torch.randn(NT, T, D, generator=g).cuda(), torch.randn(NT, T, D, generator=g).cuda(), torch.randn(NT, T, generator=g).cuda()
S = torch.empty((T, T, NT), device="cuda")
for _ in range(4):
    sip_score(q, k, d, out=S)
torch.cuda.synchronize()
print("done", float(S[5, 3, 2]))
