"""GPU (B200): the fused retrieval metrics (reference VisionLanguageModule.py:364-439; csrc/lse_fwd.cu
MODE_RANK / MODE_TOPK) against the oracle.  Ranks and neighbour indices are INTEGER-EXACT on data
whose similarities are exactly representable (small dyadic entries: fp32 tensor-core accumulation
and the fp64 oracle both compute them without rounding), including heavy exact ties from duplicated
captions; on real normalised embeddings the metrics agree with the oracle evaluated on the same
bf16-rounded operands."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import clip_oracle as O  # noqa: E402


@pytest.fixture(scope="module")
def R():
    import vlp_b200  # noqa: F401
    from vlp_b200 import _lib, retrieval
    assert os.path.exists(_lib.lib_path())
    return retrieval


def _dyadic(n, d, gen, dev, unique=None):
    """rows with entries k / 16, k in -3..3 (exact in bf16; all dot products exact in fp32 and fp64)"""
    m = n if unique is None else unique
    base = torch.randint(-3, 4, (m, d), generator=gen, device=dev).float() / 16.0
    if unique is None:
        return base
    return base[torch.randint(0, m, (n,), generator=gen, device=dev)]


def _raw_ranks(R, q_bf16, k_bf16):
    """straight through the C ABI on bf16 operands as given (no normalisation)"""
    import ctypes  # noqa: F401
    from vlp_b200 import _lib
    lib = _lib.load()
    n_rows, d = q_bf16.shape
    n_cols = k_bf16.shape[0]
    rank = torch.empty(n_rows, dtype=torch.int32, device=q_bf16.device)
    nb = lib.vlpclip_retrieval_workspace_bytes(n_rows, n_cols, d, 0)
    ws = torch.empty(max(nb, 256), dtype=torch.uint8, device=q_bf16.device)
    rc = lib.vlpclip_retrieval_ranks(q_bf16.data_ptr(), q_bf16.stride(0), k_bf16.data_ptr(), k_bf16.stride(0),
                                     n_rows, n_cols, d, rank.data_ptr(), ws.data_ptr(), nb,
                                     torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "retrieval_ranks")
    return rank


def _raw_topk(R, q_bf16, k_bf16, k):
    from vlp_b200 import _lib
    lib = _lib.load()
    n_rows, d = q_bf16.shape
    n_cols = k_bf16.shape[0]
    idx = torch.empty(n_rows, k, dtype=torch.int32, device=q_bf16.device)
    nb = lib.vlpclip_retrieval_workspace_bytes(n_rows, n_cols, d, k)
    ws = torch.empty(max(nb, 256), dtype=torch.uint8, device=q_bf16.device)
    rc = lib.vlpclip_retrieval_topk(q_bf16.data_ptr(), q_bf16.stride(0), k_bf16.data_ptr(), k_bf16.stride(0),
                                    n_rows, n_cols, d, k, idx.data_ptr(), None, ws.data_ptr(), nb,
                                    torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "retrieval_topk")
    return idx


@pytest.mark.parametrize("n,d,unique", [(30080, 64, 880), (5000, 128, None), (300, 72, 40), (129, 8, None)])
def test_ranks_and_topk_are_integer_exact_including_ties(R, n, d, unique):
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(n + d)
    img = _dyadic(n, d, gen, dev)
    txt = _dyadic(n, d, gen, dev, unique=unique)      # e.g. 880 unique captions for 30080 images: exact ties
    qb, kb = img.to(torch.bfloat16), txt.to(torch.bfloat16)
    assert torch.equal(qb.float(), img) and torch.equal(kb.float(), txt)
    got_r = _raw_ranks(R, qb, kb)
    got_t = _raw_topk(R, qb, qb, 16 if n > 16 else n)
    torch.cuda.synchronize()
    ref_r = O.retrieval_ranks(img, txt)
    ref_t = O.retrieval_topk(img, img, got_t.shape[1])
    assert torch.equal(got_r.long(), ref_r)
    assert torch.equal(got_t.long(), ref_t)


def test_metrics_on_normalised_embeddings_match_the_oracle(R):
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(5)
    n, d = 4096, 128
    img = torch.randn(n, d, generator=gen, device=dev)
    txt = img + 0.8 * torch.randn(n, d, generator=gen, device=dev)
    labels = torch.randint(0, 2, (n,), generator=gen, device=dev)
    ks = [3, 5, 10, 15]
    rec = R.recall_at_k_on_image_text_retrieval(img, txt, [1] + ks)
    pre = R.precision_at_k_on_image_embeddings(img, labels, ks)
    # the oracle on the operands the kernel ranks (bf16-rounded normalised embeddings, fp64 products)
    qb = torch.nn.functional.normalize(img).to(torch.bfloat16).double()
    kb = torch.nn.functional.normalize(txt).to(torch.bfloat16).double()
    rank = O.retrieval_ranks(qb, kb)
    top = O.retrieval_topk(qb, qb, 16)
    hits = labels[:, None] == labels[top[:, 1:]]
    for k in [1] + ks:
        assert abs(rec[k] - int((rank < k).sum()) / n) <= 2.0 / n       # fp32 vs fp64 near-ties
    for k in ks:
        assert abs(pre[k] - (hits[:, :k].sum(dim=1).float() / k).mean().item()) <= 1e-3
    # and the reference's own definition on the unrounded embeddings (statistically the same numbers)
    ref_rec = O.recall_at_k_on_image_text_retrieval(img.cpu(), txt.cpu(), [1] + ks)
    ref_pre = O.precision_at_k_on_image_embeddings(img.cpu(), labels.cpu(), ks)
    for k in [1] + ks:
        assert abs(rec[k] - ref_rec[k]) < 0.01
    for k in ks:
        assert abs(pre[k] - ref_pre[k]) < 0.01


def test_retrieval_argument_errors(R):
    dev = torch.device("cuda:0")
    e = torch.randn(64, 32, device=dev)
    with pytest.raises(ValueError):
        R.retrieval_topk(e, e, 17)
    with pytest.raises(ValueError):
        R.retrieval_ranks(e, e[:32])          # every query needs its paired key
