"""CPU: the drop-in module's host logic (constructor contract, parameter names/dtypes, optimiser
parameter groups, error behaviour, Hydra-style instantiation, retrieval metrics) -- no kernels."""
import functools
import importlib
import math
import os

import numpy as np
import pytest
import torch

os.environ.setdefault("VLP_B200_RANDOM_INIT", "1")   # no network: build encoders from configs

import vlp_b200  # noqa: E402
from vlp_b200 import functional as VF  # noqa: E402
from vlp_b200.module import LogitsHandle, VisionLanguageModule  # noqa: E402
from oracle import clip_oracle as O  # noqa: E402

REF_MODEL_CFG = "/root/reference/configs/model/vision_language.yaml"
REF_EXPERIMENT_CFG = "/root/reference/configs/experiment/pretrain/pretrain_resnet34_tinybert.yaml"


def make_module(**over):
    kw = dict(image_model="resnet18", text_encoder_model="tinybert",
              optimizer=functools.partial(torch.optim.AdamW, lr=5e-5), deduplicate=False,
              masked_loss=False, image_embedding_dim=512, text_embedding_dim=312, embedding_dim=128)
    kw.update(over)
    return VisionLanguageModule(**kw)


@pytest.fixture(scope="module")
def module():
    return make_module()


def test_parameters_follow_reference_contract(module):
    assert tuple(module.image_projection.shape) == (512, 128)
    assert tuple(module.text_projection.shape) == (312, 128)
    assert module.image_projection.dtype == torch.float32
    assert tuple(module.logit_scale.shape) == (1,)
    assert module.logit_scale.dtype == torch.float64          # reference :111 (numpy scalar -> fp64)
    assert abs(float(module.logit_scale.detach()) - math.log(1 / 0.07)) < 1e-12
    # CLIP init: std = F^-0.5 (reference :102-109)
    assert abs(module.image_projection.std().item() - 512 ** -0.5) < 0.2 * 512 ** -0.5


def test_state_dict_keys_match_reference_checkpoints(module):
    keys = set(module.state_dict().keys())
    for k in ("image_projection", "text_projection", "logit_scale",
              "image_encoder.model.conv1.weight", "image_encoder.model.layer4.1.bn2.weight"):
        assert k in keys, k
    assert any(k.startswith("text_encoder.model.") for k in keys)
    # what OnlyImagingModule / FusionModule strip to load the pretrained image encoder
    assert all(not k.startswith("image_encoder.") or k.startswith("image_encoder.model.") for k in keys)


def test_hparams_saved(module):
    assert module.hparams["embedding_dim"] == 128
    assert module.hparams["deduplicate"] is False and module.hparams["masked_loss"] is False


def test_deduplicate_overrides_masked_loss():
    m = make_module(deduplicate=True, masked_loss=True)
    assert m.hparams["masked_loss"] is False and m.hparams["deduplicate"] is True


def test_unsupported_text_encoder_raises():
    with pytest.raises(ValueError, match="is not supported"):
        make_module(text_encoder_model="roberta")


def test_optimizer_groups_default(module):
    cfg = module.configure_optimizers()
    names = [g["name"] for g in cfg["optimizer"].param_groups]
    assert names == ["remaining_params", "projection_and_logitscale", "image_encoder", "text_encoder"]
    head = cfg["optimizer"].param_groups[1]["params"]
    assert len(head) == 3
    assert module.hparams["num_optimized_params"] == sum(p.numel() for p in module.parameters())


def test_optimizer_group_lr_zero_freezes_and_drops_group():
    m = make_module(text_encoder_lr=0.0, image_encoder_lr=1e-5, projections_lr=1e-4)
    cfg = m.configure_optimizers()
    groups = {g["name"]: g for g in cfg["optimizer"].param_groups}
    assert "text_encoder" not in groups
    assert all(not p.requires_grad for p in m.text_encoder.parameters())
    assert groups["image_encoder"]["lr"] == 1e-5 and groups["projection_and_logitscale"]["lr"] == 1e-4


def test_optimizer_group_negative_lr_raises():
    m = make_module(projections_lr=-1.0)
    with pytest.raises(ValueError, match="non-negative"):
        m.configure_optimizers()


def test_scheduler_partial_is_wired():
    m = make_module(scheduler=functools.partial(torch.optim.lr_scheduler.StepLR, step_size=3))
    cfg = m.configure_optimizers()
    assert cfg["lr_scheduler"]["interval"] == "epoch" and cfg["lr_scheduler"]["frequency"] == 1


def test_deprecated_loss_flags_raise_like_reference(module):
    h = LogitsHandle(torch.zeros(2, 4), torch.zeros(2, 4), module.logit_scale)
    with pytest.raises(DeprecationWarning, match="Deduplication loss was made obsolete"):
        module._compute_loss(h)                               # default deduplicate=True (reference :532)
    with pytest.raises(DeprecationWarning, match="Masked loss was made obsolete"):
        module._compute_loss(h, deduplicate=False, masked=True)
    with pytest.raises(TypeError):
        module._compute_loss(torch.zeros(2, 2), deduplicate=False, masked=False)


def test_no_cpu_fallback():
    I, T = O.make_embeddings(8, 16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        VF.fused_clip_loss_from_embeddings(I, T, torch.tensor([2.0]))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        VF.project_normalize(torch.randn(4, 8), torch.randn(8, 8))


def test_validation_dataloader_index_guard(module):
    with pytest.raises(Exception):
        module.validation_step({"x-ray": torch.zeros(1, 3, 8, 8)}, 0, dataloader_idx=2)


def test_logits_handle_behaves_like_a_square_matrix(module):
    I, T = O.make_embeddings(6, 16)
    h = LogitsHandle(I, T, torch.tensor([2.0], dtype=torch.float64))
    assert len(h) == 6 and h.shape == (6, 6)
    ref = (I @ T.T) * torch.clamp(torch.tensor([2.0], dtype=torch.float64).exp(), max=100)   # :456-459
    np.testing.assert_allclose(h.materialize().numpy(), ref.numpy(), rtol=1e-12, atol=1e-12)


def _instantiate(cfg: dict, **extra):
    """What hydra.utils.instantiate does for this config (reference src/train.py:109-116)."""
    cfg = dict(cfg)
    target = cfg.pop("_target_")
    mod, cls = target.rsplit(".", 1)
    return getattr(importlib.import_module(mod), cls)(**cfg, **extra)


def test_hydra_style_instantiation_with_target_override():
    import yaml
    if os.path.exists(REF_MODEL_CFG):
        cfg = yaml.safe_load(open(REF_MODEL_CFG))
        assert cfg["_target_"] == "src.models.pretrain.VisionLanguageModule.VisionLanguageModule"
        exp = yaml.safe_load(open(REF_EXPERIMENT_CFG))["model"]
        cfg.update({k: v for k, v in exp.items() if not (isinstance(v, str) and "${" in v)})
    else:  # the keys of configs/model/vision_language.yaml + the experiment override
        cfg = {"_target_": "x", "image_model": "resnet34", "text_encoder_model": "distilbert",
               "deduplicate": False, "masked_loss": False, "downstream_datamodule": "downstream",
               "embedding_dim": 128, "image_embedding_dim": 512, "text_embedding_dim": 312}
    cfg["_target_"] = "vlp_b200.VisionLanguageModule"          # the drop-in switch
    cfg["text_encoder_model"] = "tinybert"                     # ${text_encoder_model} of the experiment
    cfg["image_model"] = "resnet18"                            # keep the CPU test light
    cfg["downstream_datamodule"] = None
    m = _instantiate(cfg, optimizer=functools.partial(torch.optim.AdamW, lr=5e-5), scheduler=None,
                     label_weights=(1.0, 1.0))
    assert type(m).__name__ == "VisionLanguageModule"
    assert tuple(m.text_projection.shape) == (312, 128)


def test_retrieval_metrics_have_no_cpu_path():
    """The fused retrieval metrics (csrc/lse_fwd.cu, MODE_RANK / MODE_TOPK) only run on a B200: CPU
    tensors raise instead of silently taking a torch path (GPU parity: tests/test_gpu_retrieval.py)."""
    from vlp_b200.retrieval import precision_at_k_on_image_embeddings, recall_at_k_on_image_text_retrieval
    g = torch.Generator().manual_seed(0)
    img = torch.randn(300, 32, generator=g)
    labels = torch.randint(0, 2, (300,), generator=g)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        recall_at_k_on_image_text_retrieval(img, img, [1, 5])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        precision_at_k_on_image_embeddings(img, labels, [3])
    with pytest.raises(AssertionError):      # reference :382
        precision_at_k_on_image_embeddings(img[:10], labels[:10], [15])


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` needs no GPU: one JSON line on stdout with the driver's keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--n", "2048", "--cpu-block", "128"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "pairs/s" and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["value"] > 0 and line["config"]["global_batch"] == 2048


def test_epoch_cache_appends_without_reconcatenation():
    """cache.EpochEmbeddingCache: reserve / commit semantics on CPU tensors (the buffers are plain
    torch tensors; only the prologue kernel that fills the reserved rows needs a GPU)."""
    from vlp_b200.cache import EpochEmbeddingCache
    c = EpochEmbeddingCache(initial_rows=4)
    with pytest.raises(ValueError):
        c.get()
    g = torch.Generator().manual_seed(0)
    chunks = [(torch.randn(n, 8, generator=g), torch.randn(n, 8, generator=g), torch.arange(n)) for n in (3, 5, 2)]
    # step 1: rows written "by the kernel" into the reserved views
    iv, tv = c.reserve(3, 8, "cpu")
    iv.copy_(chunks[0][0]); tv.copy_(chunks[0][1])
    c.commit(chunks[0][0], chunks[0][1], chunks[0][2], written=True)
    ptr = c.img.data_ptr()
    # step 2 outgrows the buffer (4 rows): it doubles and keeps the rows; plain copy path
    c.commit(chunks[1][0], chunks[1][1], chunks[1][2])
    assert c.img.shape[0] == 8 and c.img.data_ptr() != ptr
    # a reserve whose rows are NOT the ones committed falls back to the copy
    c.reserve(7, 8, "cpu")
    c.commit(chunks[2][0], chunks[2][1], chunks[2][2], written=True)
    i, t, l = c.get()
    assert len(c) == 10 and i.dtype == torch.bfloat16
    assert torch.equal(i.float(), torch.cat([x[0] for x in chunks]).to(torch.bfloat16).float())
    assert torch.equal(t.float(), torch.cat([x[1] for x in chunks]).to(torch.bfloat16).float())
    assert torch.equal(l, torch.cat([x[2] for x in chunks]))
    assert "label" in c and torch.equal(c["label"], l)
    c.reset()
    assert len(c) == 0 and c.img.shape[0] == 16      # buffers are kept across epochs
