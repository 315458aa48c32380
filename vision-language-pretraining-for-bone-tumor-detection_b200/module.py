"""Drop-in ``VisionLanguageModule`` whose CLIP head runs on the fused sm_100a kernels.

Boundary being replaced (reference ``src/models/pretrain/VisionLanguageModule.py``):

* ctor signature / hyper-parameters                      lines 64-128
* parameters ``image_projection`` [F_i, D], ``text_projection`` [F_t, D], ``logit_scale`` [1] fp64,
  sub-modules ``image_encoder.model`` / ``text_encoder.model``   lines 27-60, 98-111 (checkpoint keys)
* ``forward(batch) -> (logits, image_embeddings, text_embeddings)``          lines 441-461
* ``_compute_loss(logits, deduplicate, masked, captions) -> (loss, image_loss, text_loss)``  532-554
* optimiser parameter groups with per-group lr / freezing                    lines 130-297
* Lightning hooks, epoch caches, retrieval metrics                           lines 299-439, 556-705

Selection is a Hydra override: ``model._target_=vlp_b200.VisionLanguageModule`` (the reference's
``configs/model/vision_language.yaml:1`` names the class to instantiate); nothing else in
``src/train.py`` changes.  The N x N ``logits`` tensor of the reference is replaced by a
``LogitsHandle`` that only ever flows into ``_compute_loss`` (exactly how the reference uses it,
lines 635-638 / 665-668); the encoders stay stock PyTorch.

lightning / timm / hydra are optional: without lightning the class derives from a small
``nn.Module`` shim that offers ``log``, ``save_hyperparameters``, ``hparams`` and ``device``.
"""
from __future__ import annotations

import logging
from itertools import chain
from typing import Dict, Sequence

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as VF
from .cache import EpochEmbeddingCache

logger = logging.getLogger("project")

try:  # pragma: no cover - depends on the environment
    import lightning as _L
    _LightningBase = _L.LightningModule
    HAVE_LIGHTNING = True
except Exception:  # lightning is not installed in the build image
    _L = None
    HAVE_LIGHTNING = False

    class _AttrDict(dict):
        __getattr__ = dict.__getitem__
        __setattr__ = dict.__setitem__

    class _LightningBase(nn.Module):
        """Minimal stand-in for ``lightning.LightningModule`` (only what this module uses)."""

        def __init__(self):
            super().__init__()
            self._hparams = _AttrDict()
            self.logged: Dict[str, object] = {}
            self.trainer = None

        @property
        def hparams(self):
            return self._hparams

        def save_hyperparameters(self, *args, logger=True, **kwargs):
            import inspect
            frame = inspect.currentframe().f_back
            init_args = {}
            local_vars = frame.f_locals
            cls = type(self)
            sig = inspect.signature(cls.__init__)
            for name, prm in sig.parameters.items():
                if name == "self":
                    continue
                if prm.kind is inspect.Parameter.VAR_KEYWORD:
                    init_args.update(local_vars.get(name, {}))
                elif name in local_vars:
                    init_args[name] = local_vars[name]
            self._hparams.update(init_args)

        @property
        def device(self):
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")

        def log(self, name, value, *args, **kwargs):
            self.logged[name] = value.detach() if isinstance(value, torch.Tensor) else value


class _MeanMetric:
    """Weighted running mean with the torchmetrics.MeanMetric calls the module needs."""

    def __init__(self):
        self.reset()

    def reset(self):
        self._sum = 0.0
        self._weight = 0.0

    def update(self, value, weight=1.0):
        v = float(value.detach()) if isinstance(value, torch.Tensor) else float(value)
        self._sum += v * float(weight)
        self._weight += float(weight)

    def compute(self):
        return torch.tensor(self._sum / self._weight if self._weight > 0 else float("nan"))


def _mean_metric():
    try:  # pragma: no cover
        from torchmetrics import MeanMetric
        return MeanMetric()
    except Exception:
        return _MeanMetric()


# ----------------------------------------------------------------------------------------------
# encoders (OUT OF SCOPE of the fused path: stock PyTorch, only wrapped so state-dict keys match)
# ----------------------------------------------------------------------------------------------
class ImageEncoder(nn.Module):
    """timm ``create_model(model, pretrained=False, num_classes=0, global_pool='avg')`` when timm is
    importable (reference lines 27-35); otherwise the torchvision ResNet of the same name with
    ``fc = Identity`` (same parameter names: conv1, bn1, layer1..4)."""

    def __init__(self, model, **kwargs):
        super().__init__()
        self.drop_rate = float(kwargs.get("drop_rate", 0.0) or 0.0)
        try:  # pragma: no cover
            import timm
            self.model = timm.create_model(model, pretrained=False, num_classes=0,
                                           global_pool="avg", **kwargs)
            self._timm = True
        except ImportError:
            import torchvision
            if not hasattr(torchvision.models, model):
                raise ValueError(f"image model {model!r} is not available without timm")
            net = getattr(torchvision.models, model)(weights=None)
            net.fc = nn.Identity()
            self.model = net
            self._timm = False

    def forward(self, x):
        y = self.model(x)
        if not self._timm and self.drop_rate > 0.0:
            y = F.dropout(y, p=self.drop_rate, training=self.training)
        return y


class TextEncoder(nn.Module):
    """DistilBERT / TinyBERT, CLS token (reference lines 38-60).  Like the reference, construction
    fails when the pretrained weights cannot be loaded; only an explicit VLP_B200_RANDOM_INIT=1
    (benchmarks, tests: no network) builds the same architecture from its config with random
    weights."""

    def __init__(self, text_encoder_model):
        super().__init__()
        if text_encoder_model == "distilbert":
            self.model = self._load("distilbert-base-uncased", "distilbert")
        elif text_encoder_model == "tinybert":
            self.model = self._load("huawei-noah/TinyBERT_General_4L_312D", "tinybert")
        else:
            raise ValueError(
                f"VisionLanguageModule: Text encoder model {text_encoder_model} is not supported. "
                "Supported models are: distilbert, tinybert.")
        self.target_token_idx = 0
        self.model.train()

    @staticmethod
    def _load(hub_name, kind):
        import os
        import transformers
        if os.environ.get("VLP_B200_RANDOM_INIT") != "1":
            # (a failed download must not silently pre-train a random text tower: let it raise)
            if kind == "distilbert":
                return transformers.DistilBertModel.from_pretrained(hub_name)
            return transformers.AutoModel.from_pretrained(hub_name, torch_dtype="auto")
        logger.warning("TextEncoder: VLP_B200_RANDOM_INIT=1 -- building %s from its config with random weights",
                       hub_name)
        if kind == "distilbert":
            return transformers.DistilBertModel(transformers.DistilBertConfig())
        cfg = transformers.BertConfig(hidden_size=312, num_hidden_layers=4, num_attention_heads=12,
                                      intermediate_size=1200, vocab_size=30522)
        return transformers.BertModel(cfg)

    def forward(self, **kwargs):
        return self.model(**kwargs).last_hidden_state[:, self.target_token_idx, :]


# ----------------------------------------------------------------------------------------------
# the logits stand-in
# ----------------------------------------------------------------------------------------------
class LogitsHandle:
    """What ``forward`` returns instead of the N x N logit matrix.

    Holds the embeddings (and their bf16/fp16 operand copies) plus ``logit_scale``; the fused loss
    consumes it.  ``len(handle)`` is the batch size, ``handle.materialize()`` builds the real
    matrix with torch ops for debugging small batches only."""

    def __init__(self, image_embeddings, text_embeddings, logit_scale, operands=None):
        self.image_embeddings = image_embeddings
        self.text_embeddings = text_embeddings
        self.logit_scale = logit_scale
        self.operands = operands

    def __len__(self):
        return self.image_embeddings.shape[0]

    @property
    def shape(self):
        n = len(self)
        return (n, n)

    @property
    def device(self):
        return self.image_embeddings.device

    def materialize(self) -> torch.Tensor:
        s = torch.clamp(self.logit_scale.exp(), max=VF.LOGIT_SCALE_MAX)
        return (self.image_embeddings @ self.text_embeddings.T) * s


# ----------------------------------------------------------------------------------------------
# the module
# ----------------------------------------------------------------------------------------------
class VisionLanguageModule(_LightningBase):
    def __init__(
        self,
        image_model,
        text_encoder_model,
        optimizer,
        deduplicate: bool,
        masked_loss: bool,
        image_embedding_dim: int = 512,
        text_embedding_dim: int = 768,
        embedding_dim: int = 256,
        label_weights: tuple = (1.0, 1.0),   # interface parity with the classifiers; unused here
        scheduler=None,
        downstream_datamodule=None,
        text_encoder_lr: float = None,
        image_encoder_lr: float = None,
        projections_lr: float = None,
        image_encoder_droupout: float = 0.0,   # (sic) spelling kept: it is a public kwarg
        **kwargs,
    ):
        super().__init__()
        if deduplicate:
            if masked_loss:
                logger.warning("Deduplication and masked loss are mutually exclusive. "
                               "Deduplication will be used.")
            masked_loss = False
        self.save_hyperparameters(logger=False)

        # fused-head options (extra kwargs, absent from the reference's signature)
        self.share_negatives_across_ranks = bool(kwargs.get("share_negatives_across_ranks", True))
        # mask_duplicate_captions=True: samples of a batch that carry the same caption string are not
        # used as negatives of one another (fused duplicate-caption mask).  The
        # reference's deduplicate / masked_loss flags keep raising like the reference.
        self.mask_duplicate_captions = bool(kwargs.get("mask_duplicate_captions", False))

        self.image_encoder = ImageEncoder(image_model, drop_rate=image_encoder_droupout)
        self.text_encoder = TextEncoder(text_encoder_model)

        # CLIP-style init (reference lines 102-111): N(0, F^-0.5), log-temperature ln(1/0.07) in fp64
        self.image_projection = nn.Parameter(torch.empty(image_embedding_dim, embedding_dim))
        nn.init.normal_(self.image_projection, std=image_embedding_dim ** -0.5)
        self.text_projection = nn.Parameter(torch.empty(text_embedding_dim, embedding_dim))
        nn.init.normal_(self.text_projection, std=text_embedding_dim ** -0.5)
        self.logit_scale = nn.Parameter(torch.tensor([np.log(1 / 0.07)]))

        self.deduplicated_loss_function = torch.nn.BCEWithLogitsLoss()
        self.k_for_precision_at_k = [3, 5, 10, 15]
        self.k_for_image_text_retreival = [3, 5, 10, 15]
        self.val_combined_loss = _mean_metric()

        self.downstream_datamodule = downstream_datamodule
        if self.downstream_datamodule is not None:
            dm, _ = next(self.downstream_datamodule.get_cv_splits())
            self.downstream_train_dataloader = dm.train_dataloader()
            self.downstream_val_dataloaders = dm.val_dataloader()
        # epoch caches: preallocated bf16 buffers the prologue kernel writes into (cache.py)
        self.train_image_embeddings_and_labels_cached = EpochEmbeddingCache()
        self.val_image_embeddings_and_labels_cached = EpochEmbeddingCache()
        self._cache_pending = None
        logger.info("VisionLanguageModule (fused B200 head): initialised with %s", dict(self.hparams))

    # ------------------------------------------------------------------ optimiser (lines 130-297)
    def configure_optimizers(self):
        groups = self._configure_optimizer_parameters()
        optimizer = self.hparams.optimizer(params=groups)
        for g in optimizer.param_groups:
            logger.info("Parameter group '%s': %d params, lr=%s", g.get("name", "unnamed"),
                        sum(p.numel() for p in g["params"]), g.get("lr", "default"))
        n_opt = sum(p.numel() for g in optimizer.param_groups for p in g["params"])
        logger.info("VisionLanguageModule: Number of parameters optimized by the optimizer: %d", n_opt)
        self.hparams["num_optimized_params"] = n_opt
        if self.hparams.scheduler is not None:
            scheduler = self.hparams.scheduler(optimizer=optimizer)
            return {"optimizer": optimizer,
                    "lr_scheduler": {"scheduler": scheduler, "interval": "epoch", "frequency": 1}}
        return {"optimizer": optimizer}

    def _configure_optimizer_parameters(self):
        img = list(self.image_encoder.parameters())
        txt = list(self.text_encoder.parameters())
        head = [self.image_projection, self.text_projection, self.logit_scale]
        taken = {id(p) for p in img + txt + head}
        rest = [p for p in self.parameters() if id(p) not in taken]
        if rest:
            logger.warning("VisionLanguageModule: There are %d parameters that are not assigned to "
                           "any group.", len(rest))
        groups = [{"params": rest, "name": "remaining_params"}]
        for params, name, lr in ((head, "projection_and_logitscale", self.hparams.projections_lr),
                                 (img, "image_encoder", self.hparams.image_encoder_lr),
                                 (txt, "text_encoder", self.hparams.text_encoder_lr)):
            g = self._get_param_group(params, name, lr)
            if g is not None:
                groups.append(g)
        return groups

    def _get_param_group(self, params, name: str, lr):
        group = {"params": params, "name": name}
        if lr is None:
            return group
        if lr < 0:
            logger.error("VisionLanguageModule: %s scale learning rate is set to a negative value.", name)
            raise ValueError(f"VisionLanguageModule: {name} scale learning rate must be a non-negative value.")
        if lr == 0:
            for p in params:      # frozen group: not handed to the optimiser at all
                p.requires_grad = False
            return None
        group["lr"] = lr
        return group

    # ------------------------------------------------------------------ hot path
    def _process_group(self):
        if not self.share_negatives_across_ranks:
            return None, 1
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                return dist.group.WORLD, dist.get_world_size()
        except Exception:
            pass
        return None, 1

    def forward(self, batch):
        image_features = self.image_encoder(batch["x-ray"])
        text_features = self.text_encoder(**batch["caption_tokenized"])
        # projection + L2-normalise (reference :448-453) on the fused tf32 kernel; its bf16 operand
        # copy of the embeddings lands directly in the next rows of the epoch cache
        cache = (self.train_image_embeddings_and_labels_cached if self.training
                 else self.val_image_embeddings_and_labels_cached)
        i_view, t_view = cache.reserve(image_features.shape[0], self.image_projection.shape[1],
                                       image_features.device)
        i_emb, i_bf16, i_f16 = VF.project_normalize(image_features, self.image_projection, out_bf16=i_view)
        t_emb, t_bf16, t_f16 = VF.project_normalize(text_features, self.text_projection, out_bf16=t_view)
        self._cache_pending = (cache, i_emb, t_emb)
        handle = LogitsHandle(i_emb, t_emb, self.logit_scale, (i_bf16, t_bf16, i_f16, t_f16))
        return handle, i_emb, t_emb

    def _compute_loss(self, logits, deduplicate: bool = True, masked: bool = False, captions: list = None):
        if deduplicate:
            raise DeprecationWarning(
                "Deduplication loss was made obsolete by generating diverse captions and the custom batch sampler")
        if masked:
            raise DeprecationWarning(
                "Masked loss was made obsolete by generating diverse captions and the custom batch sampler")
        if not isinstance(logits, LogitsHandle):
            raise TypeError("the fused head never materialises the N x N logits: pass the handle "
                            "returned by forward() (LogitsHandle) to _compute_loss")
        group, world = self._process_group()
        if group is not None and not self._equal_shards(len(logits), group):
            # the row-sharded loss needs the same number of pairs on every rank (a remainder batch
            # of the reference's samplers may differ): this step falls back to the reference's own
            # DDP semantics, the loss over the local pairs
            group, world = None, 1
        caption_ids = None
        if self.mask_duplicate_captions and captions is not None:
            caption_ids = self._caption_ids(captions, logits.device)
        loss, image_loss, text_loss = VF.fused_clip_loss_from_embeddings(
            logits.image_embeddings, logits.text_embeddings, logits.logit_scale, group=group,
            grad_scale=float(world), caption_ids=caption_ids, _operands=logits.operands)
        return loss, image_loss, text_loss

    @staticmethod
    def _caption_ids(captions, device) -> torch.Tensor:
        """Caption strings -> non-negative int32 ids that agree across ranks without communication
        (crc32 of the text; the reference numbers them with np.unique per batch, :482)."""
        import zlib
        ids = [zlib.crc32(str(c).encode("utf-8")) & 0x7FFFFFFF for c in captions]
        return torch.tensor(ids, dtype=torch.int32, device=device)

    @staticmethod
    def _equal_shards(n_loc: int, group) -> bool:
        """True when every rank of ``group`` holds ``n_loc`` pairs (one tiny all-reduce per step)."""
        import torch.distributed as dist
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" \
            else torch.device("cpu")
        t = torch.tensor([n_loc, -n_loc], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        hi, neg_lo = t.tolist()
        if hi != -neg_lo:
            logger.warning("fused CLIP loss: ranks hold between %d and %d pairs this step; using the "
                           "local (per-rank) loss for it", -neg_lo, hi)
            return False
        return True

    # ------------------------------------------------------------------ retrieval metrics
    def precision_at_k_on_image_embeddings(self, image_embeddings, labels, ks: Sequence[int]) -> dict:
        from .retrieval import precision_at_k_on_image_embeddings
        return precision_at_k_on_image_embeddings(image_embeddings, labels, ks)

    def recall_at_k_on_image_text_retreival(self, image_embeddings, text_embeddings, ks: Sequence[int]) -> dict:
        from .retrieval import recall_at_k_on_image_text_retrieval
        return recall_at_k_on_image_text_retrieval(image_embeddings, text_embeddings, ks)

    def evaluate_downstream_precision_at_k(self, mode="entire"):
        if mode == "entire":
            batches = chain(self.downstream_train_dataloader, *self.downstream_val_dataloaders)
        elif mode == "validation":
            batches = chain(*self.downstream_val_dataloaders)
        else:
            raise ValueError(f"Invalid mode: {mode}. Supported modes are: 'entire', 'validation'.")
        embs, labels = [], []
        was_training = self.training
        self.eval()
        with torch.no_grad():
            for batch in batches:
                x = batch["x-ray"].to(device=self.device)
                y = batch["tumor"].to(device=self.device, dtype=torch.int64)
                embs.append(self.image_encoder(x) @ self.image_projection)
                labels.append(y)
        self.train(was_training)
        return self.precision_at_k_on_image_embeddings(torch.cat(embs), torch.cat(labels),
                                                       ks=self.k_for_precision_at_k)

    # ------------------------------------------------------------------ epoch caches (lines 556-628)
    def _cache_embeddings_and_labels(self, image_embeddings, text_embeddings, labels, mode):
        assert mode in ["train", "val"], f"Invalid mode: {mode}"
        cache = (self.train_image_embeddings_and_labels_cached if mode == "train"
                 else self.val_image_embeddings_and_labels_cached)
        # the embeddings forward() just produced already sit in the cache rows it reserved (written
        # by the prologue kernel); anything else is copied in (detached: no autograd graph is kept)
        pend = self._cache_pending
        written = (pend is not None and pend[0] is cache and pend[1] is image_embeddings
                   and pend[2] is text_embeddings)
        cache.commit(image_embeddings, text_embeddings, labels, written=written)
        self._cache_pending = None

    def _get_cached_embeddings_and_labels(self, mode):
        assert mode in ["train", "val"], f"Invalid mode: {mode}"
        cache = (self.train_image_embeddings_and_labels_cached if mode == "train"
                 else self.val_image_embeddings_and_labels_cached)
        if len(cache) == 0:
            raise ValueError(f"No cached embeddings and labels for mode: {mode}")
        return cache.get()      # views of the bf16 buffers: no concatenation

    # ------------------------------------------------------------------ Lightning hooks (631-705)
    def on_train_epoch_start(self):
        self.train_image_embeddings_and_labels_cached.reset()

    def training_step(self, batch, batch_idx=None):
        logits, image_embeddings, text_embeddings = self(batch)
        self._cache_embeddings_and_labels(image_embeddings, text_embeddings, batch["label"], mode="train")
        loss, _, _ = self._compute_loss(logits, self.hparams["deduplicate"], self.hparams["masked_loss"],
                                        batch.get("caption"))
        bs = batch["x-ray"].shape[0]
        self.log("train/loss", loss, on_step=True, on_epoch=True, batch_size=bs)
        self.log("logit_scale", self.logit_scale.exp(), on_step=True, on_epoch=True, batch_size=bs)
        return loss

    def on_train_epoch_end(self):
        img, txt, labels = self._get_cached_embeddings_and_labels(mode="train")
        for k, v in self.precision_at_k_on_image_embeddings(img, labels, ks=self.k_for_precision_at_k).items():
            self.log(f"train/label_precision_at_{k}", v, on_step=False, on_epoch=True, batch_size=img.shape[0])
        for k, v in self.recall_at_k_on_image_text_retreival(img, txt, ks=self.k_for_image_text_retreival).items():
            self.log(f"train/image_text_recall_at_{k}", v, on_step=False, on_epoch=True, batch_size=img.shape[0])

    def on_validation_epoch_start(self):
        self.val_combined_loss.reset()
        self.val_image_embeddings_and_labels_cached.reset()

    def validation_step(self, batch, batch_idx, dataloader_idx=0):
        logits, image_embeddings, text_embeddings = self(batch)
        self._cache_embeddings_and_labels(image_embeddings, text_embeddings, batch["label"], mode="val")
        loss, _, _ = self._compute_loss(logits, self.hparams["deduplicate"], self.hparams["masked_loss"],
                                        batch.get("caption"))
        if dataloader_idx == 0:
            prefix = "val/lera"
        elif dataloader_idx == 1:
            prefix = "val/mura"
        else:
            raise ValueError(
                f"VisionLanguageModule: Validation dataloader index {dataloader_idx} is not supported. "
                "Supported indices are: 0, 1. We are assuming that the first dataloader is for the LERA "
                "dataset and the second dataloader for the MURA dataset")
        bs = batch["x-ray"].shape[0]
        self.log(f"{prefix}/loss", loss, on_step=False, on_epoch=True, batch_size=bs, add_dataloader_idx=False)
        self.val_combined_loss.update(loss.detach(), bs)
        return loss

    def on_validation_epoch_end(self):
        self.log("val/combined/loss", self.val_combined_loss.compute(), prog_bar=True)
        img, txt, labels = self._get_cached_embeddings_and_labels(mode="val")
        for k, v in self.precision_at_k_on_image_embeddings(img, labels, ks=self.k_for_precision_at_k).items():
            self.log(f"val/combined/label_precision_at_{k}", v, batch_size=img.shape[0])
        for k, v in self.recall_at_k_on_image_text_retreival(img, txt, ks=self.k_for_image_text_retreival).items():
            self.log(f"val/combined/image_text_recall_at_{k}", v, batch_size=img.shape[0])
        trainer = getattr(self, "trainer", None)
        if trainer is not None and getattr(trainer, "sanity_checking", False):
            logger.info("VisionLanguageModule: Skipping downstream zero-shot evaluation during sanity check.")
            return
        if self.downstream_datamodule is None:
            return
        res = self.evaluate_downstream_precision_at_k(mode="validation")
        if res:
            for k, v in res.items():
                self.log(f"downstream_validation/label_precision_at_{k}", v, on_step=False, on_epoch=True)
