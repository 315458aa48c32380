"""GPU (B200): projection + normalise prologue, the full head and the drop-in module."""
import functools
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
os.environ.setdefault("VLP_B200_RANDOM_INIT", "1")

from oracle import clip_oracle as O  # noqa: E402


@pytest.fixture(scope="module")
def VF():
    import vlp_b200  # noqa: F401
    from vlp_b200 import functional
    return functional


@pytest.mark.parametrize("n,fi,ft,d", [(256, 512, 312, 512), (300, 512, 768, 128), (1024, 2048, 312, 256), (64, 512, 312, 32),
                                       (384, 512, 312, 768)])
def test_head_against_straight_through_oracle(VF, n, fi, ft, d):
    dev = torch.device("cuda:0")
    ls = math.log(1 / 0.07)
    f_i, f_t, w_i, w_t = O.make_features(n, fi, ft, d, seed=7)
    fic, ftc, wic, wtc = (t.to(dev).requires_grad_(True) for t in (f_i, f_t, w_i, w_t))
    lsc = torch.tensor([ls], dtype=torch.float64, device=dev, requires_grad=True)
    loss, il, tl, ie, te = VF.fused_clip_loss(fic, ftc, wic, wtc, lsc)
    loss.backward()
    torch.cuda.synchronize()
    # (1) embeddings: within one bf16 ulp (2^-8 relative to the largest component) of fp64 normalize
    Ei = torch.nn.functional.normalize(f_i.double() @ w_i.double())
    Et = torch.nn.functional.normalize(f_t.double() @ w_t.double())
    assert (ie.detach().cpu().double() - Ei).abs().max().item() < 2.0 ** -9 * Ei.abs().max().item() * 2
    assert (te.detach().cpu().double() - Et).abs().max().item() < 2.0 ** -9 * Et.abs().max().item() * 2
    # (2) loss on the kernel's own bf16 embeddings: tolerance 1e-4 relative
    ib = ie.detach().to(torch.bfloat16).float().cpu()
    tb = te.detach().to(torch.bfloat16).float().cpu()
    ref = O.closed_form(ib.numpy(), tb.numpy(), ls)
    assert abs(loss.item() - ref["loss"]) < 1e-4 * ref["loss"]
    # (3) gradients through normalise + projection (fp64), tolerance 1e-3 normwise
    ui = (f_i.double() @ w_i.double()).numpy()
    ut = (f_t.double() @ w_t.double()).numpy()
    dui = O.normalize_backward(ui, ref["dI"])
    dut = O.normalize_backward(ut, ref["dT"])
    assert O.rel_err(wic.grad.cpu().numpy(), f_i.double().numpy().T @ dui) < 1e-3
    assert O.rel_err(wtc.grad.cpu().numpy(), f_t.double().numpy().T @ dut) < 1e-3
    assert O.rel_err(fic.grad.cpu().numpy(), dui @ w_i.double().numpy().T) < 1e-3
    assert O.rel_err(ftc.grad.cpu().numpy(), dut @ w_t.double().numpy().T) < 1e-3
    assert abs(lsc.grad.item() - ref["dlogit_scale"]) < 1e-3 * abs(ref["dlogit_scale"])


def test_head_against_reference_golden(VF, golden_dir):
    """Reference forward + _compute_loss run in fp32 (golden) vs the fused head: the loss differs
    only by the bf16 rounding of the embeddings (<= 3e-4 relative)."""
    g = dict(np.load(os.path.join(golden_dir, "head_n32_f512_312_d128_fp32.npz")))
    n, f_img, f_txt, d, ls, seed, _ = g["params"]
    fi, ft, wi, wt = O.make_features(int(n), int(f_img), int(f_txt), int(d), seed=int(seed))
    dev = torch.device("cuda:0")
    loss, il, tl, ie, te = VF.fused_clip_loss(fi.to(dev), ft.to(dev), wi.to(dev), wt.to(dev),
                                              torch.tensor([ls], dtype=torch.float64, device=dev))
    assert np.abs(ie.cpu().numpy() - g["image_embeddings"]).max() < 1e-3
    assert np.abs(te.cpu().numpy() - g["text_embeddings"]).max() < 1e-3
    assert abs(loss.item() - float(np.ravel(g["loss"])[0])) < 1e-3 * float(np.ravel(g["loss"])[0])


def test_module_training_step_end_to_end():
    import vlp_b200  # noqa: F401
    from vlp_b200.module import LogitsHandle, VisionLanguageModule
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    m = VisionLanguageModule(image_model="resnet18", text_encoder_model="tinybert",
                             optimizer=functools.partial(torch.optim.AdamW, lr=5e-5), deduplicate=False,
                             masked_loss=False, image_embedding_dim=512, text_embedding_dim=312,
                             embedding_dim=128).to(dev)
    bsz = 24
    batch = {"x-ray": torch.randn(bsz, 1, 64, 64, device=dev).repeat(1, 3, 1, 1),
             "caption_tokenized": {"input_ids": torch.randint(0, 30522, (bsz, 16), device=dev),
                                   "token_type_ids": torch.zeros(bsz, 16, dtype=torch.long, device=dev),
                                   "attention_mask": torch.ones(bsz, 16, dtype=torch.long, device=dev)},
             "label": torch.randint(0, 2, (bsz,), device=dev), "caption": ["c"] * bsz}
    opt = m.configure_optimizers()["optimizer"]
    m.on_train_epoch_start()
    loss = m.training_step(batch)
    assert torch.isfinite(loss)
    loss.backward()
    for name in ("image_projection", "text_projection", "logit_scale"):
        g = getattr(m, name).grad
        assert g is not None and torch.isfinite(g).all() and g.abs().sum() > 0
    assert m.logit_scale.grad.dtype == torch.float64
    assert any(p.grad is not None and p.grad.abs().sum() > 0 for p in m.image_encoder.parameters())
    opt.step()
    # the handle path and the reference formula agree on the module's own embeddings
    with torch.no_grad():
        handle, ie, te = m(batch)
        assert isinstance(handle, LogitsHandle)
        l2, _, _ = m._compute_loss(handle, False, False, None)
        ref = O.closed_form(ie.to(torch.bfloat16).float().cpu().numpy(), te.to(torch.bfloat16).float().cpu().numpy(),
                            float(m.logit_scale.detach()))
        assert abs(l2.item() - ref["loss"]) < 1e-4 * ref["loss"]
    m.on_train_epoch_end()
    m.on_validation_epoch_start()
    m.validation_step(batch, 0, dataloader_idx=0)
    m.on_validation_epoch_end()
    assert "val/combined/loss" in getattr(m, "logged", {"val/combined/loss": 1})
