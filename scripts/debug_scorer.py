"""GPU diagnostics for the tcgen05 scorer: exact-integer inputs, prints where the output differs."""
import sys
import time
import torch
sys.path.insert(0, ".")
from transkun_b200.LayersTransformer import sip_score

def run(NT, T, D, seed=0, timing=False):
    g = torch.Generator().manual_seed(seed)
    q = torch.randint(-3, 4, (NT, T, D), generator=g).float()
    k = torch.randint(-3, 4, (NT, T, D), generator=g).float()
    diag = torch.randn(NT, T, generator=g)
    qc, kc, dc = q.cuda(), k.cuda(), diag.cuda()
    S = sip_score(qc, kc, dc)
    torch.cuda.synchronize()
    t = torch.arange(T, dtype=torch.float32)
    want = (torch.einsum("ned,nbd->neb", q, k) / (D ** 0.5)) * (t[:, None] - t[None, :]).abs()
    want = (want + torch.diag_embed(diag)).permute(1, 2, 0)
    tri = torch.tril(torch.ones(T, T, dtype=torch.bool))
    bad = ((S.cpu() != want) & tri[:, :, None])
    msg = f"NT={NT} T={T} D={D}: bad {int(bad.sum())} / {int(tri.sum()) * NT}"
    if bad.any():
        idx = bad.nonzero()
        e, b, n = idx[0].tolist()
        msg += f" first (e={e},b={b},n={n}) got {float(S[e,b,n])} want {float(want[e,b,n])}; bad e range {int(idx[:,0].min())}-{int(idx[:,0].max())} b range {int(idx[:,1].min())}-{int(idx[:,1].max())} tracks {sorted(set(idx[:,2].tolist()))[:10]}"
    if timing:
        for _ in range(3): sip_score(qc, kc, dc, out=S)
        torch.cuda.synchronize(); t0 = time.time()
        for _ in range(10): sip_score(qc, kc, dc, out=S)
        torch.cuda.synchronize(); dt = (time.time() - t0) / 10
        ob = 4 * NT * T * (T + 1) / 2
        msg += f" | {dt*1e6:.0f} us, out {ob/dt/1e9:.0f} GB/s, {2*D*NT*T*(T+1)/2/dt/1e12:.1f} TFLOP/s"
    print(("OK   " if not bad.any() else "FAIL ") + msg, flush=True)

if __name__ == "__main__":
    for a in [(8, 64, 32), (8, 128, 32), (8, 128, 256), (8, 200, 256), (3, 70, 64), (16, 257, 256), (90, 691, 256)]:
        run(*a)
    run(88, 2048, 256, timing=True)
