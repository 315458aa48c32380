"""GPU (B200): parity of the fused CUDA path against the oracle / the reference's golden vectors.

Tolerances are the north-star's: loss within 1e-4 relative, gradients within 1e-3 (normwise
relative, ||g - g_ref|| / ||g_ref||) versus the fp32/fp64 reference on the SAME bf16-representable
embeddings.  Everything goes through the C ABI (ctypes) via vlp_b200.functional."""
import glob
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import clip_oracle as O  # noqa: E402

LOSS_RTOL = 1e-4
GRAD_RTOL = 1e-3


@pytest.fixture(scope="module")
def VF():
    import vlp_b200  # noqa: F401
    from vlp_b200 import _lib, functional
    assert os.path.exists(_lib.lib_path())
    return functional


def run_fused(VF, I, T, ls, ls_dtype=torch.float64):
    dev = torch.device("cuda:0")
    Ic = I.to(dev).requires_grad_(True)
    Tc = T.to(dev).requires_grad_(True)
    lsc = torch.tensor([ls], dtype=ls_dtype, device=dev, requires_grad=True)
    loss, il, tl = VF.fused_clip_loss_from_embeddings(Ic, Tc, lsc)
    loss.backward()
    torch.cuda.synchronize()
    return dict(loss=loss.item(), image_loss=il.item(), text_loss=tl.item(),
                dI=Ic.grad.cpu().numpy(), dT=Tc.grad.cpu().numpy(), dl=lsc.grad.item())


def assert_close(got, ref, ref_dl):
    for k in ("loss", "image_loss", "text_loss"):
        assert abs(got[k] - ref[k]) <= LOSS_RTOL * abs(ref[k]) + 1e-9, (k, got[k], ref[k])
    assert O.rel_err(got["dI"], ref["dI"]) < GRAD_RTOL
    assert O.rel_err(got["dT"], ref["dT"]) < GRAD_RTOL
    if ref_dl == 0.0:
        assert got["dl"] == 0.0
    else:
        assert abs(got["dl"] - ref_dl) <= GRAD_RTOL * abs(ref_dl) + 1e-12


EMB_CASES = sorted(os.path.basename(p)[:-4] for p in
                   glob.glob(os.path.join(os.path.dirname(__file__), "golden", "emb_*.npz")))


@pytest.mark.parametrize("name", EMB_CASES)
def test_against_reference_golden_vectors(VF, golden_dir, name):
    """Outputs of the reference's own forward/_compute_loss code (fp32) on seeded embeddings."""
    g = dict(np.load(os.path.join(golden_dir, name + ".npz")))
    n, d, rho, ls, seed = g["params"]
    I, T = O.make_embeddings(int(n), int(d), rho=float(rho), seed=int(seed))
    got = run_fused(VF, I, T, float(ls))
    ref = {k: float(np.ravel(g[k])[0]) for k in ("loss", "image_loss", "text_loss")}
    ref["dI"], ref["dT"] = g["dI"], g["dT"]
    assert_close(got, ref, float(np.ravel(g["d_logit_scale"])[0]))


@pytest.mark.parametrize("n,d,ls,rho", [
    (256, 512, math.log(1 / 0.07), 0.35),   # BASELINE config 2
    (256, 512, math.log(1 / 0.07), 0.0),
    (256, 512, math.log(50.0), 0.35),       # tiny loss: positive pair dominates
    (256, 512, math.log(50.0), 0.0),
    (256, 512, 5.0, 0.0),                   # clamped: s = 100, d logit_scale = 0
    (1, 64, 2.0, 0.0),                      # single pair: loss 0, zero gradients
    (2, 8, 2.6593, 0.35),                   # smallest supported embedding dim
    (100, 72, 3.0, 0.35),                   # ragged rows / cols / K
    (129, 40, 2.6593, 0.2),                 # one row past a tile boundary
    (1000, 128, 3.0, 0.5),
    (2048, 256, 2.6593, 0.35),
    (4096, 512, 2.6593, 0.35),              # BASELINE config 3 size on one GPU
    (4096, 512, 4.6, 0.35),
    (256, 768, 2.6593, 0.35),               # BASELINE config 5 embedding dim: single S buffer, 2-pass dX
    (1000, 640, 3.0, 0.35),
    (2048, 768, 2.6593, 0.35),
])
def test_against_closed_form_oracle(VF, n, d, ls, rho):
    I, T = O.make_embeddings(n, d, rho=rho, seed=42)
    ref = O.closed_form(I.numpy(), T.numpy(), ls)
    got = run_fused(VF, I, T, ls)
    assert_close(got, ref, ref["dlogit_scale"])


def test_logit_scale_dtype_and_shape_follow_the_parameter(VF):
    I, T = O.make_embeddings(64, 64, seed=1)
    dev = torch.device("cuda:0")
    for dt in (torch.float64, torch.float32):
        ls = torch.tensor([2.5], dtype=dt, device=dev, requires_grad=True)
        loss, _, _ = VF.fused_clip_loss_from_embeddings(I.to(dev).requires_grad_(True), T.to(dev), ls)
        loss.backward()
        assert ls.grad.dtype == dt and tuple(ls.grad.shape) == (1,)
        assert loss.dtype == torch.float32


def test_forward_only_no_grad_mode(VF):
    I, T = O.make_embeddings(300, 128, seed=2)
    dev = torch.device("cuda:0")
    with torch.no_grad():
        loss, il, tl = VF.fused_clip_loss_from_embeddings(I.to(dev), T.to(dev), torch.tensor([3.0], device=dev))
    ref = O.closed_form(I.numpy(), T.numpy(), 3.0)
    assert abs(loss.item() - ref["loss"]) < LOSS_RTOL * ref["loss"]
    assert not loss.requires_grad


def test_needs_input_grad_is_honoured(VF):
    """Frozen groups (reference :273-281): any of dI / dT / d logit_scale may be not required."""
    I, T = O.make_embeddings(200, 64, seed=3)
    ref = O.closed_form(I.numpy(), T.numpy(), 2.6593)
    dev = torch.device("cuda:0")
    Ic = I.to(dev).requires_grad_(True)
    ls = torch.tensor([2.6593], dtype=torch.float64, device=dev)          # frozen temperature
    loss, _, _ = VF.fused_clip_loss_from_embeddings(Ic, T.to(dev), ls)     # frozen text tower
    loss.backward()
    assert O.rel_err(Ic.grad.cpu().numpy(), ref["dI"]) < GRAD_RTOL
    ls2 = torch.tensor([2.6593], dtype=torch.float64, device=dev, requires_grad=True)
    loss2, _, _ = VF.fused_clip_loss_from_embeddings(I.to(dev), T.to(dev), ls2)   # only the temperature
    loss2.backward()
    assert abs(ls2.grad.item() - ref["dlogit_scale"]) < GRAD_RTOL * abs(ref["dlogit_scale"])


def test_image_and_text_loss_backpropagate_separately(VF):
    I, T = O.make_embeddings(150, 64, seed=4)
    dev = torch.device("cuda:0")
    Ic = I.to(dev).requires_grad_(True)
    Tc = T.to(dev).requires_grad_(True)
    ls = torch.tensor([2.6593], dtype=torch.float64, device=dev)
    _, il, _ = VF.fused_clip_loss_from_embeddings(Ic, Tc, ls)
    il.backward()
    Ir = I.double().clone().requires_grad_(True)
    Tr = T.double().clone().requires_grad_(True)
    _, il_ref, _ = O.loss_from_embeddings(Ir, Tr, torch.tensor([2.6593], dtype=torch.float64))
    il_ref.backward()
    assert O.rel_err(Ic.grad.cpu().numpy(), Ir.grad.numpy()) < GRAD_RTOL
    assert O.rel_err(Tc.grad.cpu().numpy(), Tr.grad.numpy()) < GRAD_RTOL


def test_bitwise_reproducible(VF):
    I, T = O.make_embeddings(1500, 256, seed=5)
    a = run_fused(VF, I, T, 2.6593)
    b = run_fused(VF, I, T, 2.6593)
    assert a["loss"] == b["loss"] and a["dl"] == b["dl"]
    assert np.array_equal(a["dI"], b["dI"]) and np.array_equal(a["dT"], b["dT"])


@pytest.mark.parametrize("n,d", [(1000, 128), (3000, 768), (5000, 512)])
def test_single_sweep_backward_matches_two_pass(VF, n, d):
    """vlpclip_grad_both (one sweep feeds dI and dT) against two vlpclip_grad passes: same G tiles,
    so the results agree to fp32 summation order; both are bit-reproducible."""
    dev = torch.device("cuda:0")
    I, T = O.make_embeddings(n, d, rho=0.35, seed=11)
    Ib, Tb = I.to(dev).to(torch.bfloat16), T.to(dev).to(torch.bfloat16)
    s = 1 / 0.07
    rm, rl, rdiag, cm, cl = VF.lse_stats_fused(Ib, Tb, s, 0)
    r, c = VF.merge_stats(rm, rl, rdiag, s)[:3], VF.merge_stats(cm, cl, rdiag, s)[:3]
    i16, t16 = VF.cast_bf16_to_f16(Ib), VF.cast_bf16_to_f16(Tb)
    dI2, ds2 = VF._grad(i16, t16, r, c, s, 0, n, 1.0, 1.0, True)
    dT2, _ = VF._grad(t16, i16, c, r, s, 0, n, 1.0, 1.0, False)
    dI, dT, ds = VF._grad_both(i16, t16, r, c, s, 0, n, 1.0, 1.0, True)
    dIb, dTb, dsb = VF._grad_both(i16, t16, r, c, s, 0, n, 1.0, 1.0, True)
    torch.cuda.synchronize()
    assert torch.equal(dI, dIb) and torch.equal(dT, dTb) and torch.equal(ds, dsb)
    rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm()).item()  # noqa: E731
    assert rel(dI, dI2) < 2e-5 and rel(dT, dT2) < 2e-5
    assert abs(ds.item() - ds2.item()) <= 1e-5 * abs(ds2.item())


def test_single_sweep_backward_bf16_outputs_and_upstream_gradient(VF):
    """bf16 gradient outputs take the partial-buffer path of the dT consumers (pieces collect in fp32,
    the last piece reads them back); an upstream gradient is folded into the epilogues."""
    dev = torch.device("cuda:0")
    n, d = 3000, 256
    I, T = O.make_embeddings(n, d, rho=0.35, seed=12)
    Ib, Tb = I.to(dev).to(torch.bfloat16), T.to(dev).to(torch.bfloat16)
    s = 1 / 0.07
    rm, rl, rdiag, cm, cl = VF.lse_stats_fused(Ib, Tb, s, 0)
    r, c = VF.merge_stats(rm, rl, rdiag, s)[:3], VF.merge_stats(cm, cl, rdiag, s)[:3]
    i16, t16 = VF.cast_bf16_to_f16(Ib), VF.cast_bf16_to_f16(Tb)
    mul = torch.tensor([0.37], dtype=torch.float32, device=dev)
    dI, dT, _ = VF._grad_both(i16, t16, r, c, s, 0, n, 1.0, 1.0, False)
    dIh, dTh, _ = VF._grad_both(i16, t16, r, c, s, 0, n, 1.0, 1.0, False, mul,
                                (torch.bfloat16, torch.bfloat16))
    torch.cuda.synchronize()
    assert dIh.dtype == torch.bfloat16 and dTh.dtype == torch.bfloat16
    rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm()).item()  # noqa: E731
    assert rel(dIh, 0.37 * dI) < 4e-3 and rel(dTh, 0.37 * dT) < 4e-3      # bf16 rounding of the outputs


@pytest.mark.parametrize("n,d,n_captions,ls", [(300, 64, 40, 2.6593), (1000, 128, 88, 3.0), (4096, 512, 880, 2.6593),
                                               (129, 72, 5, math.log(50.0))])
def test_duplicate_caption_mask_matches_oracle(VF, n, d, n_captions, ls):
    """f3: pairs (i, j != i) with the same caption id are excluded from both cross-entropies
    (reference _get_mask, :506-530); loss within 1e-4, gradients within 1e-3 of the fp64 oracle."""
    dev = torch.device("cuda:0")
    I, T = O.make_embeddings(n, d, rho=0.35, seed=21)
    g = torch.Generator().manual_seed(5)
    ids = torch.randint(0, n_captions, (n,), generator=g)
    Ic = I.to(dev).requires_grad_(True)
    Tc = T.to(dev).requires_grad_(True)
    lsc = torch.tensor([ls], dtype=torch.float64, device=dev, requires_grad=True)
    loss, il, tl = VF.fused_clip_loss_from_embeddings(Ic, Tc, lsc, caption_ids=ids.to(dev))
    loss.backward()
    torch.cuda.synchronize()
    ref = O.masked_loss_and_grads_from_embeddings(I, T, torch.tensor([ls], dtype=torch.float64), ids)
    unmasked = O.closed_form(I.numpy(), T.numpy(), ls)
    assert abs(float(ref["loss"]) - unmasked["loss"]) > 2e-4 * unmasked["loss"]      # the mask matters here
    for k, v in (("loss", loss), ("image_loss", il), ("text_loss", tl)):
        assert abs(v.item() - float(ref[k])) <= LOSS_RTOL * abs(float(ref[k])), k
    assert O.rel_err(Ic.grad.cpu().numpy(), ref["dI"].numpy()) < GRAD_RTOL
    assert O.rel_err(Tc.grad.cpu().numpy(), ref["dT"].numpy()) < GRAD_RTOL
    assert abs(lsc.grad.item() - float(ref["dlogit_scale"])) <= GRAD_RTOL * abs(float(ref["dlogit_scale"]))
    # all captions distinct: exactly the unmasked loss
    Ic.grad = Tc.grad = lsc.grad = None
    loss_u, _, _ = VF.fused_clip_loss_from_embeddings(Ic, Tc, lsc, caption_ids=torch.arange(n, device=dev))
    plain, _, _ = VF.fused_clip_loss_from_embeddings(Ic, Tc, lsc)
    assert loss_u.item() == plain.item()


def test_errors_are_raised_not_swallowed(VF):
    dev = torch.device("cuda:0")
    I, T = O.make_embeddings(16, 16)
    with pytest.raises(ValueError):
        VF.fused_clip_loss_from_embeddings(I.to(dev), T[:8].to(dev), torch.tensor([2.0], device=dev))
    with pytest.raises(ValueError):   # embedding dim must be a multiple of 8
        VF.fused_clip_loss_from_embeddings(torch.randn(16, 12, device=dev), torch.randn(16, 12, device=dev),
                                           torch.tensor([2.0], device=dev))
    with pytest.raises(ValueError):   # > 768 columns do not fit the TMEM plan
        VF.fused_clip_loss_from_embeddings(torch.randn(16, 1024, device=dev), torch.randn(16, 1024, device=dev),
                                           torch.tensor([2.0], device=dev))


# ---- full size (BASELINE headline: 32768 x 512): size-independent properties -------------------
@pytest.fixture(scope="module")
def full_size(VF):
    n, d, ls = 32768, 512, math.log(1 / 0.07)
    I, T = O.make_embeddings(n, d, rho=0.35, seed=42)
    got = run_fused(VF, I, T, ls)
    return n, d, ls, I, T, got


def test_full_size_sampled_rows_against_fp64(full_size):
    n, d, ls, I, T, got = full_size
    s = math.exp(ls)
    rows = np.random.RandomState(0).choice(n, 48, replace=False)
    Id, Td = I.double(), T.double()
    S_rows = s * (Id[rows] @ Td.T)                      # [48, n]
    S_cols = s * (Td[rows] @ Id.T)                      # rows of S^T
    lse_r = torch.logsumexp(S_rows, dim=1)
    lse_c_sel = torch.logsumexp(S_cols, dim=1)
    # dI_i = s/(2N) * sum_j (P_row_ij + P_col_ij - 2 delta_ij) T_j needs every column LSE
    lse_c_all = torch.cat([torch.logsumexp(s * (Td[lo:lo + 2048] @ Id.T), dim=1) for lo in range(0, n, 2048)])
    G = torch.exp(S_rows - lse_r[:, None]) + torch.exp(S_rows - lse_c_all[None, :])
    G[torch.arange(len(rows)), torch.as_tensor(rows)] -= 2.0
    dI_ref = (s / (2.0 * n)) * (G @ Td)
    assert O.rel_err(got["dI"][rows], dI_ref.numpy()) < GRAD_RTOL
    # loss terms of the sampled rows are consistent with the global mean (sanity on magnitude)
    loss_rows = 0.5 * ((lse_r - torch.diagonal(S_rows[:, rows])) + (lse_c_sel - torch.diagonal(S_cols[:, rows])))
    assert abs(loss_rows.mean().item() - got["loss"]) < 0.25 * got["loss"]


def test_full_size_temperature_checksum(full_size):
    """sum_i <I_i, dI_i> = sum_j <T_j, dT_j> = d logit_scale (unclamped), a checksum of checksums."""
    n, d, ls, I, T, got = full_size
    a = float((I.double().numpy() * got["dI"].astype(np.float64)).sum())
    b = float((T.double().numpy() * got["dT"].astype(np.float64)).sum())
    assert abs(a - got["dl"]) < 2e-3 * abs(got["dl"])
    assert abs(b - got["dl"]) < 2e-3 * abs(got["dl"])


def test_full_size_loss_against_blocked_fp64(full_size):
    n, d, ls, I, T, got = full_size
    s = math.exp(ls)
    Id, Td = I.double(), T.double()
    il = tl = 0.0
    for lo in range(0, n, 2048):
        Sr = s * (Id[lo:lo + 2048] @ Td.T)
        Sc = s * (Td[lo:lo + 2048] @ Id.T)
        idx = torch.arange(lo, min(n, lo + 2048))
        il += float((torch.logsumexp(Sr, 1) - Sr[torch.arange(len(idx)), idx]).sum())
        tl += float((torch.logsumexp(Sc, 1) - Sc[torch.arange(len(idx)), idx]).sum())
    ref = 0.5 * (il + tl) / n
    assert abs(got["loss"] - ref) < LOSS_RTOL * ref


def test_permutation_equivariance(VF):
    I, T = O.make_embeddings(1024, 128, seed=6)
    perm = torch.randperm(1024, generator=torch.Generator().manual_seed(1))
    a = run_fused(VF, I, T, 2.6593)
    b = run_fused(VF, I[perm], T[perm], 2.6593)
    assert abs(a["loss"] - b["loss"]) < 1e-5 * a["loss"]
    assert O.rel_err(b["dI"], a["dI"][perm.numpy()]) < 2e-4


def test_wide_lse_spread_takes_the_two_exp_path(VF):
    """Half of the pairs perfectly aligned (LSE ~ s), the other half unrelated (LSE ~ 0.3 s) at the
    clamped temperature: the global log2-LSE spread exceeds the single-ex2 guard of the backward,
    which must fall back to two ex2 per logit and stay within tolerance."""
    n, d, ls = 384, 64, 5.0
    g = torch.Generator().manual_seed(11)
    I = torch.nn.functional.normalize(torch.randn(n, d, generator=g)).to(torch.bfloat16).float()
    T = I.clone()
    T[n // 2:] = torch.nn.functional.normalize(torch.randn(n - n // 2, d, generator=g)).to(torch.bfloat16).float()
    ref = O.closed_form(I.numpy(), T.numpy(), ls)
    spread = (ref["row_lse"].max() - ref["row_lse"].min()) * math.log2(math.e)
    assert spread > 60.0
    got = run_fused(VF, I, T, ls)
    assert_close(got, ref, ref["dlogit_scale"])


def test_fused_and_exact_column_statistics_agree(VF):
    """vlpclip_lse_fwd_fused (one sweep) vs two vlpclip_lse_fwd sweeps, after the merge."""
    dev = torch.device("cuda:0")
    for (n_rows, n_cols, d, s, shift) in [(256, 256, 512, 14.29, 0), (300, 1000, 72, 50.0, 0),
                                          (1024, 4096, 256, 100.0, -1024), (4096, 4096, 512, 100.0, 0)]:
        n = max(n_rows, n_cols)
        I, T = O.make_embeddings(n, d, rho=0.35, seed=3)
        x = I[:n_rows].to(dev).to(torch.bfloat16).contiguous()
        y = T[:n_cols].to(dev).to(torch.bfloat16).contiguous()
        rm, rl, rd, cm, cl = VF.lse_stats_fused(x, y, s, shift)
        rm2, rl2, rd2 = VF.lse_stats(x, y, s, shift)
        cm2, cl2, cd2 = VF.lse_stats(y, x, s, -shift)
        assert torch.equal(rm, rm2) and torch.equal(rd, rd2)
        assert torch.allclose(rl, rl2, rtol=1e-6, atol=0)
        # compare column LSEs (natural log) -- references differ, the totals must not
        cdiag = torch.zeros(n_cols, device=dev)
        idx = torch.arange(n_rows, device=dev) - shift
        ok = (idx >= 0) & (idx < n_cols)
        cdiag[idx[ok]] = rd[ok]
        lse_f = VF.merge_stats(cm, cl, cdiag, s, want_lse=True)[4]
        lse_e = VF.merge_stats(cm2, cl2, cd2, s, want_lse=True)[4]
        has = torch.zeros(n_cols, dtype=torch.bool, device=dev)
        has[idx[ok]] = True
        assert torch.allclose(lse_f[has], lse_e[has], rtol=2e-6, atol=2e-5)


def test_cuda_graph_path_matches_eager(VF, monkeypatch):
    """fwd+bwd captured as one CUDA graph (default for sharded runs) vs the eager path."""
    I, T = O.make_embeddings(700, 128, seed=8)
    eager = run_fused(VF, I, T, 2.6593)
    monkeypatch.setattr(VF, "_GRAPH_MODE", "1")
    VF._GRAPHS.clear()
    outs = [run_fused(VF, I, T, 2.6593) for _ in range(4)]     # call 1 eager warm-up, 2 captures, 3+ replay
    I2, T2 = O.make_embeddings(700, 128, seed=9)                # new data through the same graph
    other = run_fused(VF, I2, T2, 3.0)
    monkeypatch.setattr(VF, "_GRAPH_MODE", "0")
    ref2 = O.closed_form(I2.numpy(), T2.numpy(), 3.0)
    for o in outs:
        assert o["loss"] == eager["loss"] and o["dl"] == eager["dl"]
        assert np.array_equal(o["dI"], eager["dI"]) and np.array_equal(o["dT"], eager["dT"])
    assert_close(other, ref2, ref2["dlogit_scale"])
    VF._GRAPHS.clear()


def test_largest_sweep_size_properties(VF):
    """BASELINE config 5 upper end (65536 x 512): size-independent properties only -- the
    temperature checksum sum_i <I_i, dI_i> = sum_j <T_j, dT_j> = d logit_scale, run-to-run
    bit-reproducibility, and the loss of rows whose statistics are recomputed in fp64."""
    n, d, ls = 65536, 512, math.log(1 / 0.07)
    I, T = O.make_embeddings(n, d, rho=0.35, seed=7)
    a = run_fused(VF, I, T, ls)
    b = run_fused(VF, I, T, ls)
    assert a["loss"] == b["loss"] and np.array_equal(a["dI"], b["dI"]) and np.array_equal(a["dT"], b["dT"])
    ci = float((I.double().numpy() * a["dI"].astype(np.float64)).sum())
    ct = float((T.double().numpy() * a["dT"].astype(np.float64)).sum())
    assert abs(ci - a["dl"]) < 2e-3 * abs(a["dl"]) and abs(ct - a["dl"]) < 2e-3 * abs(a["dl"])
    s = math.exp(ls)
    rows = np.random.RandomState(1).choice(n, 32, replace=False)
    S_rows = s * (I.double()[rows] @ T.double().T)
    img_rows = torch.logsumexp(S_rows, 1) - S_rows[torch.arange(32), torch.as_tensor(rows)]
    S_cols = s * (T.double()[rows] @ I.double().T)
    txt_rows = torch.logsumexp(S_cols, 1) - S_cols[torch.arange(32), torch.as_tensor(rows)]
    # the sampled per-pair losses bracket the global mean
    est = 0.5 * (img_rows.mean().item() + txt_rows.mean().item())
    assert abs(est - a["loss"]) < 0.25 * a["loss"]


# ---- small entry points added for the sharded / graphed step -----------------------------------
DEV = torch.device("cuda:0")


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("ls", [math.log(1 / 0.07), math.log(50.0), math.log(100.0) - 1e-3, 5.0, -3.0])
def test_scale_prep_matches_reference_exp_clamp(VF, dtype, ls):
    """reference :456-457: s = clamp(exp(l), max=100); ds/dl = exp(l) below the clamp, else 0."""
    l = torch.tensor([ls], dtype=dtype, device=DEV, requires_grad=True)
    s_ref = torch.clamp(l.exp(), max=100)
    s_ref.backward()
    s, ds = VF.scale_from_logit_scale(l)
    assert s.dtype == torch.float32 and ds.dtype == torch.float32
    assert abs(s.item() - s_ref.item()) <= 1e-6 * abs(s_ref.item())
    assert abs(ds.item() - l.grad.item()) <= 1e-6 * max(abs(l.grad.item()), 1e-30)


def test_slot_sum_is_an_ordered_sum(VF):
    from vlp_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(3)
    slots = torch.randn(5, 300, 72, generator=g).to(DEV)
    mul = torch.tensor([0.37], device=DEV)
    want = slots[0].clone()
    for s in range(1, 5):
        want += slots[s]
    want = want * mul
    for out_dtype in (torch.float32, torch.bfloat16):
        out = torch.empty(300, 72, dtype=out_dtype, device=DEV)
        rc = lib.vlpclip_slot_sum(slots.data_ptr(), 5, 300 * 72, mul.data_ptr(),
                                  1 if out_dtype == torch.bfloat16 else 0, out.data_ptr(),
                                  torch.cuda.current_stream().cuda_stream)
        assert rc == 0, lib.vlpclip_last_error()
        torch.cuda.synchronize()
        assert torch.equal(out, want.to(out_dtype))


@pytest.mark.parametrize("n,d,owners", [(1000, 72, 4), (4096, 512, 8), (384, 128, 3)])
def test_grad_scatter_routes_rows_to_their_owner(VF, n, d, owners):
    """The fused reduce-scatter's store side on ONE device: owner buffers are slices of a local
    tensor, so the scattered result must equal vlpclip_grad's rows, bit for bit."""
    import ctypes

    from vlp_b200 import _lib
    lib = _lib.load()
    I, T = O.make_embeddings(n, d, rho=0.35, seed=11)
    s = math.exp(2.6593)
    i16 = VF.cast_bf16_to_f16(I.to(DEV).to(torch.bfloat16))
    t16 = VF.cast_bf16_to_f16(T.to(DEV).to(torch.bfloat16))
    rm, rl, rdiag = VF.lse_stats(I.to(DEV).to(torch.bfloat16), T.to(DEV).to(torch.bfloat16), s, 0)
    cm, cl, cdiag = VF.lse_stats(T.to(DEV).to(torch.bfloat16), I.to(DEV).to(torch.bfloat16), s, 0)
    r_stats = VF.merge_stats(rm, rl, rdiag, s)[:3]
    c_stats = VF.merge_stats(cm, cl, cdiag, s)[:3]
    want, ds_want = VF._grad(i16, t16, r_stats, c_stats, s, 0, n, 1.0, 1.0, True)
    rows_per_owner = -(-n // owners)
    # every owner gets its own (deliberately shuffled) region of one big buffer
    pool = torch.full((owners, rows_per_owner, d), float("nan"), device=DEV)
    order = list(reversed(range(owners)))
    ptrs = (ctypes.c_void_p * owners)(*[pool[order[o]].data_ptr() for o in range(owners)])
    sc = VF.as_scale_tensor(s, i16.device)
    ds = torch.zeros(1, device=DEV)
    nbytes = lib.vlpclip_grad_workspace_bytes(n, n, d)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
    rc = lib.vlpclip_grad_scatter(
        i16.data_ptr(), i16.stride(0), t16.data_ptr(), t16.stride(0), r_stats[0].data_ptr(),
        r_stats[1].data_ptr(), r_stats[2].data_ptr(), c_stats[0].data_ptr(), c_stats[1].data_ptr(),
        c_stats[2].data_ptr(), n, n, d, sc.data_ptr(), 0, n, 1.0, 1.0, ptrs, owners, rows_per_owner,
        ds.data_ptr(), ws.data_ptr(), nbytes, torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.vlpclip_last_error()
    torch.cuda.synchronize()
    for o in range(owners):
        lo, hi = o * rows_per_owner, min(n, (o + 1) * rows_per_owner)
        got = pool[order[o]][:hi - lo]
        assert torch.equal(got, want[lo:hi]), f"owner {o}"
        assert torch.isnan(pool[order[o]][hi - lo:]).all()        # rows past n are never written
    assert ds.item() == ds_want.item()


def test_forward_sweep_on_fp16_operands_matches_bf16(VF):
    """The sharded step sweeps the fp16 operand copies (one gathered copy of T serves forward and
    backward): bf16-representable values convert exactly, so the statistics must agree."""
    I, T = O.make_embeddings(1500, 256, rho=0.35, seed=21)
    s = math.exp(2.6593)
    ib, tb = I.to(DEV).to(torch.bfloat16), T.to(DEV).to(torch.bfloat16)
    ih, th = VF.cast_bf16_to_f16(ib), VF.cast_bf16_to_f16(tb)
    a = VF.lse_stats_fused(ib, tb, s, 0)
    b = VF.lse_stats_fused(ih, th, s, 0)
    for x, y in zip(a, b):
        assert torch.allclose(x, y, rtol=1e-6, atol=1e-7)
    a = VF.lse_stats(tb, ib, s, 0)
    b = VF.lse_stats(th, ih, s, 0)
    for x, y in zip(a, b):
        assert torch.allclose(x, y, rtol=1e-6, atol=1e-7)


def test_cast_push_writes_every_destination(VF):
    import ctypes

    from vlp_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(5)
    src = torch.randn(777, 72, generator=g).to(DEV).to(torch.bfloat16)
    want = VF.cast_bf16_to_f16(src)
    pool = torch.zeros(3, 1000, 72, dtype=torch.float16, device=DEV)
    ptrs = (ctypes.c_void_p * 3)(*[pool[r, 100:].data_ptr() for r in range(3)])
    rc = lib.vlpclip_cast_push_f16(src.data_ptr(), src.numel(), ptrs, 3,
                                   torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.vlpclip_last_error()
    torch.cuda.synchronize()
    for r in range(3):
        assert torch.equal(pool[r, 100:877], want)
        assert (pool[r, :100] == 0).all() and (pool[r, 877:] == 0).all()
