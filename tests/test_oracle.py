"""CPU: the oracle against (a) golden vectors produced by the reference's own code
(tests/golden/make_golden.py), (b) closed-form gradients, (c) analytic known answers."""
import glob
import json
import math
import os

import numpy as np
import pytest
import torch

from oracle import clip_oracle as O


def _load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name + ".npz")))


def test_golden_meta_points_at_reference_lines(golden_dir):
    meta = json.load(open(os.path.join(golden_dir, "golden_meta.json")))
    assert meta["forward_lines"] == [441, 461]
    assert meta["compute_loss_lines"] == [532, 554]
    assert "Deduplication loss was made obsolete" in meta["deprecation_10"]
    assert "Masked loss was made obsolete" in meta["deprecation_01"]


HEAD_CASES = ["head_n32_f512_312_d128_fp32", "head_n64_f128_40_d64_fp64", "head_n48_f64_40_d32_clamped"]


@pytest.mark.parametrize("name", HEAD_CASES)
def test_oracle_head_matches_reference_golden(golden_dir, name):
    g = _load(golden_dir, name)
    n, f_img, f_txt, d, ls, seed, bits = g["params"]
    dtype = torch.float64 if bits == 64 else torch.float32
    fi, ft, wi, wt = O.make_features(int(n), int(f_img), int(f_txt), int(d), seed=int(seed))
    for t, key in ((fi, "image_features"), (ft, "text_features"), (wi, "image_projection"), (wt, "text_projection")):
        v = t.to(dtype).double().numpy()
        np.testing.assert_allclose([v.sum(), np.abs(v).sum()], g[key + "_checksum"], rtol=1e-12)
    res = O.head_loss_and_grads(fi.to(dtype), ft.to(dtype), wi.to(dtype), wt.to(dtype),
                                torch.tensor([ls], dtype=torch.float64), dtype=torch.float64 if bits == 64 else torch.float32)
    tol = 1e-10 if bits == 64 else 2e-5
    # the literal restatement must reproduce the reference outputs (same torch, same ops)
    logits, ie, te = O.reference_forward(fi.to(dtype), ft.to(dtype), wi.to(dtype), wt.to(dtype),
                                         torch.tensor([ls], dtype=torch.float64))
    np.testing.assert_allclose(logits.numpy(), g["logits"], rtol=tol, atol=tol)
    np.testing.assert_allclose(ie.numpy(), g["image_embeddings"], rtol=tol, atol=tol)
    np.testing.assert_allclose(te.numpy(), g["text_embeddings"], rtol=tol, atol=tol)
    loss, il, tl = O.reference_compute_loss(logits)
    assert abs(float(loss) - float(g["loss"])) <= tol * max(1.0, abs(float(g["loss"])))
    assert abs(float(il) - float(g["image_loss"])) <= tol * max(1.0, abs(float(g["image_loss"])))
    assert abs(float(tl) - float(g["text_loss"])) <= tol * max(1.0, abs(float(g["text_loss"])))
    gtol = 1e-9 if bits == 64 else 1e-4
    for key, ref_key in (("d_image_features", "d_image_features"), ("d_text_features", "d_text_features"),
                         ("d_image_projection", "d_image_projection"), ("d_text_projection", "d_text_projection")):
        assert O.rel_err(res[key].numpy(), g[ref_key]) < gtol, key
    if float(ls) >= math.log(100.0):   # clamped: zero gradient (reference :457)
        assert float(np.ravel(g["d_logit_scale"])[0]) == 0.0 and float(np.ravel(res["dlogit_scale"].numpy())[0]) == 0.0
    else:
        assert abs(float(np.ravel(res["dlogit_scale"].numpy())[0]) - float(np.ravel(g["d_logit_scale"])[0])) <= gtol * abs(float(np.ravel(g["d_logit_scale"])[0]))


EMB_CASES = sorted(os.path.basename(p)[:-4] for p in
                   glob.glob(os.path.join(os.path.dirname(__file__), "golden", "emb_*.npz")))


@pytest.mark.parametrize("name", EMB_CASES)
def test_closed_form_matches_reference_golden(golden_dir, name):
    g = _load(golden_dir, name)
    n, d, rho, ls, seed = g["params"]
    I, T = O.make_embeddings(int(n), int(d), rho=float(rho), seed=int(seed))
    np.testing.assert_allclose([I.double().sum(), I.double().abs().sum()], g["I_checksum"], rtol=1e-12)
    cf = O.closed_form(I.numpy(), T.numpy(), float(ls))
    # reference ran in fp32 (embeddings) x fp64 (logit_scale): agree to fp32 round-off
    assert abs(cf["loss"] - float(g["loss"])) <= 2e-6 * max(1.0, abs(cf["loss"])) + 1e-7
    assert O.rel_err(g["dI"], cf["dI"]) < 5e-5
    assert O.rel_err(g["dT"], cf["dT"]) < 5e-5
    if cf["dlogit_scale"] == 0.0:
        assert float(np.ravel(g["d_logit_scale"])[0]) == 0.0
    else:
        assert abs(float(np.ravel(g["d_logit_scale"])[0]) - cf["dlogit_scale"]) <= 1e-4 * abs(cf["dlogit_scale"]) + 1e-9


@pytest.mark.parametrize("n,d,ls", [(17, 8, 1.0), (64, 32, 2.6593), (96, 16, 4.0), (33, 24, 5.0)])
def test_closed_form_equals_autograd_fp64(n, d, ls):
    I, T = O.make_embeddings(n, d, rho=0.3, seed=n, round_bf16=False)
    ref = O.loss_and_grads_from_embeddings(I, T, torch.tensor([ls], dtype=torch.float64), dtype=torch.float64)
    cf = O.closed_form(I.double().numpy(), T.double().numpy(), ls)
    assert abs(cf["loss"] - float(ref["loss"])) < 1e-12
    assert abs(cf["image_loss"] - float(ref["image_loss"])) < 1e-12
    assert O.rel_err(cf["dI"], ref["dI"].numpy()) < 1e-12
    assert O.rel_err(cf["dT"], ref["dT"].numpy()) < 1e-12
    assert abs(cf["dlogit_scale"] - float(np.ravel(ref["dlogit_scale"].numpy())[0])) < 1e-12 * max(1.0, abs(cf["dlogit_scale"]))


def test_known_answer_identical_rows():
    n, d = 37, 16
    v = torch.nn.functional.normalize(torch.randn(1, d, dtype=torch.float64))
    I = v.repeat(n, 1)
    cf = O.closed_form(I.numpy(), I.numpy(), 2.0)
    assert abs(cf["loss"] - math.log(n)) < 1e-12            # uniform softmax
    assert abs(cf["dlogit_scale"]) < 1e-12


def test_known_answer_orthonormal():
    n = 24
    I = np.eye(n)
    ls = 2.6593
    s = math.exp(ls)
    cf = O.closed_form(I, I, ls)
    assert abs(cf["loss"] - (math.log(math.exp(s) + n - 1) - s)) < 1e-9


def test_clamp_edge_zero_gradient():
    I, T = O.make_embeddings(32, 16, seed=3)
    res = O.loss_and_grads_from_embeddings(I, T, torch.tensor([5.0], dtype=torch.float64), dtype=torch.float64)
    assert float(np.ravel(res["dlogit_scale"].numpy())[0]) == 0.0               # e^5 = 148 > 100
    assert O.closed_form(I.numpy(), T.numpy(), 5.0)["scale"] == 100.0


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_decomposition_equals_single_process(world):
    I, T = O.make_embeddings(96, 32, seed=11, round_bf16=False)
    a = O.closed_form(I.numpy(), T.numpy(), 2.6593)
    b = O.sharded_closed_form(I.numpy(), T.numpy(), 2.6593, world)
    for k in ("loss", "image_loss", "text_loss", "dlogit_scale"):
        assert abs(a[k] - b[k]) < 1e-12
    assert O.rel_err(b["dI"], a["dI"]) < 1e-12 and O.rel_err(b["dT"], a["dT"]) < 1e-12


def test_normalize_backward_equals_autograd():
    u = torch.randn(19, 12, dtype=torch.float64, requires_grad=True)
    g = torch.randn(19, 12, dtype=torch.float64)
    torch.nn.functional.normalize(u).backward(g)
    assert O.rel_err(O.normalize_backward(u.detach().numpy(), g.numpy()), u.grad.numpy()) < 1e-12


def test_deprecated_flags_raise():
    with pytest.raises(DeprecationWarning):
        O.reference_compute_loss(torch.zeros(2, 2), deduplicate=True)
    with pytest.raises(DeprecationWarning):
        O.reference_compute_loss(torch.zeros(2, 2), masked=True)


def test_retrieval_oracle_perfect_alignment():
    e = torch.nn.functional.normalize(torch.randn(40, 16))
    r = O.recall_at_k_on_image_text_retrieval(e, e, [1, 3])
    assert r[1] == 1.0 and r[3] == 1.0
    labels = torch.arange(40) % 2
    p = O.precision_at_k_on_image_embeddings(e, labels, [3])
    assert 0.0 <= p[3] <= 1.0


def test_rank_and_stable_topk_formulation_equals_the_reference_topk_metrics():
    """The kernels compute ranks / a stable top-k (oracle.retrieval_ranks / retrieval_topk); on
    tie-free data that is exactly the reference's topk-based recall@k and precision@k."""
    g = torch.Generator().manual_seed(3)
    img = torch.nn.functional.normalize(torch.randn(500, 48, generator=g))
    txt = torch.nn.functional.normalize(img + 0.7 * torch.randn(500, 48, generator=g))
    labels = torch.randint(0, 3, (500,), generator=g)
    ks = [1, 3, 5, 10, 15]
    rank = O.retrieval_ranks(img, txt)
    ref_r = O.recall_at_k_on_image_text_retrieval(img, txt, ks)
    for k in ks:
        assert abs(int((rank < k).sum()) / 500 - ref_r[k]) < 1e-12
    top = O.retrieval_topk(img, img, 16)
    ref_p = O.precision_at_k_on_image_embeddings(img, labels, [3, 5, 10, 15])
    hits = labels[:, None] == labels[top[:, 1:]]
    for k in (3, 5, 10, 15):
        assert abs((hits[:, :k].sum(dim=1).float() / k).mean().item() - ref_p[k]) < 1e-6
    # ties: ascending index wins
    q = torch.tensor([[1.0, 0.0]]); kk = torch.tensor([[1.0, 0.0], [1.0, 0.0], [0.0, 1.0]])
    assert O.retrieval_topk(q, kk, 3)[0].tolist() == [0, 1, 2]
    assert O.retrieval_ranks(torch.cat([q, q]), kk).tolist() == [0, 1]


def test_duplicate_caption_mask_matches_the_reference_get_mask(golden_dir):
    """`oracle.reference_get_mask` (and the id mapping of the drop-in module) against the mask that the
    REFERENCE'S OWN `_get_mask` source (lines 506-530, executed by tests/golden/make_golden_mask.py)
    returns for a list of caption strings with duplicates."""
    g = np.load(os.path.join(golden_dir, "mask_captions.npz"))
    captions = [str(c) for c in g["captions"]]
    ref_mask = torch.from_numpy(g["mask"])
    assert tuple(int(v) for v in g["lines"]) == (506, 530)
    # ids the way the reference builds them (:520-521) ...
    uniq = {c: i for i, c in enumerate(sorted(set(captions)))}
    ids = torch.tensor([uniq[c] for c in captions])
    assert torch.equal(O.reference_get_mask(ids), ref_mask)
    # ... and the way the drop-in module builds them (crc32: consistent across ranks without communication)
    import zlib
    ids2 = torch.tensor([zlib.crc32(c.encode("utf-8")) & 0x7FFFFFFF for c in captions], dtype=torch.int32)
    assert torch.equal(O.reference_get_mask(ids2), ref_mask)
    assert int((ref_mask == 0).sum()) > 0 and bool((ref_mask.diagonal() == 1).all())


def test_retrieval_metric_oracle_matches_the_reference_source(golden_dir):
    """`oracle.recall_at_k_on_image_text_retrieval` / `precision_at_k_on_image_embeddings` against the values
    the REFERENCE'S OWN methods (lines 364-439, executed by tests/golden/make_golden_retrieval.py) return
    on seeded embeddings."""
    g = np.load(os.path.join(golden_dir, "retrieval_metrics.npz"))
    assert [int(v) for v in g["lines"]] == [364, 400, 402, 439]
    for tag in ("a", "b"):
        n, d, rho, seed = g[f"{tag}_params"]
        img, txt = O.make_embeddings(int(n), int(d), rho=float(rho), seed=int(seed))
        labels = torch.from_numpy(np.random.default_rng(int(seed)).integers(0, 7, size=int(n)))
        r = O.recall_at_k_on_image_text_retrieval(img, txt, [1, 3, 5, 10])
        p = O.precision_at_k_on_image_embeddings(img, labels, [3, 5, 10, 15])
        assert [r[k] for k in (1, 3, 5, 10)] == list(g[f"{tag}_recall"])
        assert [p[k] for k in (3, 5, 10, 15)] == list(g[f"{tag}_precision"])
