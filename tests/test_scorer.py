"""Scaled-Inner-Product interval scorer: oracle vs the reference's golden outputs (CPU) and the tcgen05
kernel vs the oracle (GPU, through the C ABI).  Tolerance: the kernel multiplies TF32 operands (10-bit
mantissa, the reference's --allow_tf32 regime) and accumulates in fp32: |err| <= 2e-3 * max|S| per case;
inputs that TF32 represents exactly (small integers) must come out bit-exact."""
import glob
import math
import os

import numpy as np
import pytest
import torch

from oracle.sip_oracle import sip_forward

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "scorer_*.npz")))


@pytest.mark.parametrize("path", GOLD, ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_matches_reference(path):
    z = np.load(path)
    S, b = sip_forward(z["ctx"], z["weight"], z["bias"])
    np.testing.assert_allclose(S, z["S"], rtol=1e-4, atol=1e-4)
    assert b.shape == z["b"].shape and not z["b"].any()


def _lower(S):
    T = S.shape[0]
    iu = np.tril_indices(T)
    return S[iu[0], iu[1]]


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=lambda p: os.path.basename(p)[:-4])
def test_kernel_matches_reference_module(path):
    from transkun_b200.LayersTransformer import ScaledInnerProductIntervalScorer
    z = np.load(path)
    size = z["ctx"].shape[-1]
    m = ScaledInnerProductIntervalScorer(size, 1).cuda()
    m.load_state_dict({"map.0.weight": torch.from_numpy(z["weight"]), "map.0.bias": torch.from_numpy(z["bias"])})
    with torch.no_grad():
        S, b = m(torch.from_numpy(z["ctx"]).cuda())
    assert S.shape == z["S"].shape and b.shape == z["b"].shape and not b.any()
    got, want = _lower(S.cpu().numpy()), _lower(z["S"])
    # default = the reference's inference regime (allow_tf32 unset): 3xTF32, fp32-grade products
    assert np.abs(got - want).max() <= 2e-5 * np.abs(want).max()
    assert not np.triu(S.cpu().numpy()[:, :, 0, 0], 1).any(), "cells above the diagonal are defined (zero)"
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True  # the reference's --allow_tf32 training regime: one TF32 pass
    try:
        with torch.no_grad():
            S1, _ = m(torch.from_numpy(z["ctx"]).cuda())
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    assert np.abs(_lower(S1.cpu().numpy()) - want).max() <= 2e-3 * np.abs(want).max()


@pytest.mark.gpu
@pytest.mark.parametrize("NT,T,D", [(8, 128, 256), (90, 257, 256), (5, 97, 64)])
def test_kernel_precise_mode_is_fp32_grade(NT, T, D):
    """3xTF32 (operands split into a TF32-exact part and a residual, contracted as [hi,hi,lo] x [hi,lo,hi]) against a
    float64 evaluation of the reference formula: error at the level of fp32 accumulation, three orders of magnitude
    below the one-pass TF32 kernel."""
    from transkun_b200.LayersTransformer import sip_score
    g = torch.Generator().manual_seed(NT + T)
    q, k, diag = torch.randn(NT, T, D, generator=g), torch.randn(NT, T, D, generator=g), torch.randn(NT, T, generator=g)
    t = torch.arange(T, dtype=torch.float64)
    want = (torch.einsum("ned,nbd->neb", q.double(), k.double()) / (D ** 0.5)) * (t[:, None] - t[None, :]).abs()
    want = (want + torch.diag_embed(diag.double())).permute(1, 2, 0)
    tri = torch.tril(torch.ones(T, T, dtype=torch.bool))
    scale = want[tri].abs().max().item()
    err_p = (sip_score(q.cuda(), k.cuda(), diag.cuda(), precise=True).cpu().double() - want)[tri].abs().max().item()
    err_1 = (sip_score(q.cuda(), k.cuda(), diag.cuda(), precise=False).cpu().double() - want)[tri].abs().max().item()
    assert err_p <= 3e-5 * scale, (err_p, scale)
    assert err_1 <= 2e-3 * scale and err_1 > 10 * err_p


@pytest.mark.gpu
@pytest.mark.parametrize("NT,T,D", [(8, 64, 32), (8, 128, 256), (3, 70, 64), (16, 200, 256), (90, 257, 256), (9, 129, 96)])
def test_kernel_exact_on_tf32_representable_inputs(NT, T, D):
    """Small-integer q/k are exact in TF32 and their dot products exact in fp32: every lower-triangle entry
    must equal the fp32 reference formula bit for bit (isolates tiling / swizzle / descriptor / mask logic)."""
    from transkun_b200.LayersTransformer import sip_score
    g = torch.Generator().manual_seed(NT * 1000 + T)
    q = torch.randint(-3, 4, (NT, T, D), generator=g).float()
    k = torch.randint(-3, 4, (NT, T, D), generator=g).float()
    diag = torch.randn(NT, T, generator=g)
    S = sip_score(q.cuda(), k.cuda(), diag.cuda()).cpu()
    t = torch.arange(T, dtype=torch.float32)
    want = (torch.einsum("ned,nbd->neb", q, k) / (D ** 0.5)) * (t[:, None] - t[None, :]).abs()
    want = (want + torch.diag_embed(diag)).permute(1, 2, 0)
    tri = torch.tril(torch.ones(T, T, dtype=torch.bool))
    if (D & (D - 1)) == 0 and int(D ** 0.5) ** 2 == D:   # 1/sqrt(D) exact
        assert torch.equal(S[tri], want[tri])
    else:
        torch.testing.assert_close(S[tri], want[tri], rtol=1e-6, atol=1e-5)


@pytest.mark.gpu
def test_kernel_random_inputs_and_crf_roundtrip():
    """Random fp32 inputs at the model's shape (T=691, 90 symbols, D=256): TF32 tolerance, and the scorer's
    output feeds the CRF directly (same layout, lower triangle is all it reads)."""
    from transkun_b200.CRF import NeuralSemiCRFInterval
    from transkun_b200.LayersTransformer import ScaledInnerProductIntervalScorer
    torch.manual_seed(0)
    m = ScaledInnerProductIntervalScorer(256, 1).cuda()
    ctx = torch.randn(1, 90, 691, 256, device="cuda") * 0.5
    with torch.no_grad():
        S, b = m(ctx)
        W, bias = m.map[0].weight.double(), m.map[0].bias.double()
        y = ctx.double() @ W.T + bias
        qd, kd, dd = y[..., :256] / 16.0, y[..., 256:512], y[..., 512]
        for p in (0, 45, 89):
            ref = (qd[0, p] @ kd[0, p].T)
            t = torch.arange(691, device="cuda", dtype=torch.float64)
            ref = ref * (t[:, None] - t[None, :]).abs() + torch.diag(dd[0, p])
            tri = torch.tril(torch.ones(691, 691, dtype=torch.bool, device="cuda"))
            err = (S[:, :, 0, p].double() - ref)[tri].abs().max()
            assert err <= 2e-3 * ref[tri].abs().max()
        crf = NeuralSemiCRFInterval(S.flatten(-2, -1), b.flatten(-2, -1))
        dec = crf.decode()
        assert len(dec) == 90 and torch.isfinite(crf.computeLogZ(noBackward=True)).all()


@pytest.mark.gpu
@pytest.mark.parametrize("B,P,T,D", [(1, 4, 50, 64), (2, 45, 70, 256), (1, 33, 97, 32)])
def test_scorer_backward_matches_torch_formula(B, P, T, D):
    """Gradients w.r.t. ctx and the projection weights through the kernel pair (forward tcgen05, backward
    tkb_sip_backward_prep + two library GEMMs) against autograd of the reference's formula; the upstream gradient is
    dense (non-zero above the diagonal too: those outputs are constants, their gradient must be ignored)."""
    from transkun_b200.LayersTransformer import ScaledInnerProductIntervalScorer
    torch.manual_seed(1)
    m = ScaledInnerProductIntervalScorer(D, 1).cuda()
    ctx = torch.randn(B, P, T, D, device="cuda", requires_grad=True)
    S, _ = m(ctx)
    w = torch.randn_like(S)
    tri = torch.tril(torch.ones(T, T, device="cuda"))[:, :, None, None]
    (S * w).sum().backward()
    g1, gw1 = ctx.grad.clone(), m.map[0].weight.grad.clone()
    ctx.grad = None
    m.zero_grad()
    y = m.map(ctx)
    q, k, d = y[..., :D] / math.sqrt(D), y[..., D:2 * D], y[..., 2 * D]
    t = torch.arange(T, device="cuda", dtype=torch.float32)
    S2 = torch.einsum("iped,ipbd->ipeb", q, k) * (t[:, None] - t[None, :]).abs() + torch.diag_embed(d)
    (S2.permute(2, 3, 0, 1) * w * tri).sum().backward()
    torch.testing.assert_close(g1, ctx.grad, rtol=2e-3, atol=2e-3 * float(ctx.grad.abs().max()))
    torch.testing.assert_close(gw1, m.map[0].weight.grad, rtol=2e-3, atol=2e-3 * float(gw1.abs().max()))


@pytest.mark.gpu
@pytest.mark.parametrize("rows,D", [(7, 32), (691, 256), (1000, 96)])
def test_split3_matches_the_definition(rows, D):
    """tkb_sip_split3 = [hi | hi | lo], [hi | lo | hi] with hi/lo of LayersTransformer._split_tf32, bit for bit."""
    from transkun_b200 import _lib
    from transkun_b200.LayersTransformer import _split_tf32
    g = torch.Generator().manual_seed(rows)
    q, k = torch.randn(rows, D, generator=g).cuda() * 3, torch.randn(rows, D, generator=g).cuda() * 0.01
    q3, k3 = torch.empty(rows, 3 * D, device="cuda"), torch.empty(rows, 3 * D, device="cuda")
    rc = _lib.load().tkb_sip_split3(q.data_ptr(), k.data_ptr(), rows, D, q3.data_ptr(), k3.data_ptr(),
                                    torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "tkb_sip_split3")
    qh, ql = _split_tf32(q)
    kh, kl = _split_tf32(k)
    assert torch.equal(q3, torch.cat([qh, qh, ql], 1)) and torch.equal(k3, torch.cat([kh, kl, kh], 1))
    assert torch.equal(qh + ql, q)


@pytest.mark.gpu
@pytest.mark.parametrize("NT,T,D,pitch", [(1, 1, 32, 1), (1, 2, 32, 3), (2, 33, 64, 2), (7, 64, 32, 7), (13, 130, 96, 14),
                                          (12, 65, 160, 12), (17, 200, 32, 20), (90, 129, 64, 96)])
def test_kernel_edges_small_and_unaligned(NT, T, D, pitch):
    """Shapes at the edges of the tiling: T below one tile, a last track group of 1..7 tracks, an odd number of K chunks
    (the second chunk of the last TMA box is out of bounds and zero-filled), and track pitches that take the 4-, 16- and
    32-byte store paths.  Integer inputs: bit-exact against the fp32 formula wherever 1/sqrt(D) is exact, and nothing is
    written outside the lower triangle or outside the first NT tracks of a cell."""
    from transkun_b200.LayersTransformer import sip_score
    g = torch.Generator().manual_seed(NT * 131 + T)
    q = torch.randint(-3, 4, (NT, T, D), generator=g).float()
    k = torch.randint(-3, 4, (NT, T, D), generator=g).float()
    diag = torch.randn(NT, T, generator=g)
    buf = torch.full((T, T, pitch), 7.0, device="cuda")
    S = sip_score(q.cuda(), k.cuda(), diag.cuda(), out=buf[:, :, :NT], precise=False)
    assert S.data_ptr() == buf.data_ptr()
    t = torch.arange(T, dtype=torch.float32)
    want = (torch.einsum("ned,nbd->neb", q, k) / (D ** 0.5)) * (t[:, None] - t[None, :]).abs()
    want = (want + torch.diag_embed(diag)).permute(1, 2, 0)
    tri = torch.tril(torch.ones(T, T, dtype=torch.bool))
    got = buf.cpu()
    torch.testing.assert_close(got[:, :, :NT][tri], want[tri], rtol=1e-6, atol=1e-5)
    assert (got[:, :, :NT][~tri] == 7.0).all(), "cells above the diagonal are not touched"
    assert (got[:, :, NT:] == 7.0).all(), "padding tracks are not touched"
