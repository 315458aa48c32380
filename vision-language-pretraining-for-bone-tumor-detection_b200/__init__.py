"""B200-native fused CLIP (symmetric InfoNCE) head -- drop-in for the hot path of
``src/models/pretrain/VisionLanguageModule.py`` of the reference.

Public surface:
    fused_clip_loss_from_embeddings, fused_clip_loss   (functional.py, torch.autograd Functions)
    VisionLanguageModule                                (module.py, Hydra ``_target_`` drop-in)
"""
__version__ = "0.1.0"

_LAZY = {
    "fused_clip_loss_from_embeddings": "functional",
    "fused_clip_loss": "functional",
    "clip_lse_stats": "functional",
    "VisionLanguageModule": "module",
    "ImageEncoder": "module",
    "TextEncoder": "module",
    "LogitsHandle": "module",
    "precision_at_k_on_image_embeddings": "retrieval",
    "recall_at_k_on_image_text_retrieval": "retrieval",
}


def __getattr__(name):
    if name in _LAZY:
        import importlib
        mod = importlib.import_module(f"{__name__}.{_LAZY[name]}")
        return getattr(mod, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
