"""Epoch cache of the embeddings the retrieval metrics run on (reference
``VisionLanguageModule._cache_embeddings_and_labels`` / ``_get_cached_embeddings_and_labels``,
``VisionLanguageModule.py:556-628``).

The reference re-concatenates the whole cache on every step (O(steps^2) bytes copied per epoch) and
caches the live, non-detached embeddings (keeping every step's autograd graph alive).  Here the
cache is a preallocated bf16 buffer per stream: ``reserve`` hands out the next ``rows`` rows, the
prologue kernel (``vlpclip_project_normalize_fwd``) writes its bf16 operand copy of the embeddings
straight into them -- the copy the loss kernels read anyway, so caching costs no extra pass over
the embeddings -- and ``commit`` records the labels and advances the cursor.  The buffers are kept
across epochs (``reset`` only rewinds the cursor) and double when an epoch outgrows them.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch


class EpochEmbeddingCache:
    def __init__(self, initial_rows: int = 4096):
        self.initial_rows = int(initial_rows)
        self.img: Optional[torch.Tensor] = None     # [capacity, d] bf16
        self.txt: Optional[torch.Tensor] = None
        self.lab: Optional[torch.Tensor] = None     # [capacity]
        self.n = 0
        self._pending = None                        # (row offset, rows) handed out by reserve()

    # ------------------------------------------------------------------ writing
    def reset(self) -> None:
        self.n = 0
        self._pending = None

    def _grow(self, need: int, d: int, device, label_like: Optional[torch.Tensor]) -> None:
        cap = 0 if self.img is None else self.img.shape[0]
        if self.img is not None and (self.img.shape[1] != d or self.img.device != device):
            self.img = self.txt = self.lab = None      # shape / device changed: start over
            cap, self.n = 0, 0
        if need <= cap:
            return
        new_cap = max(self.initial_rows, cap)
        while new_cap < need:
            new_cap *= 2
        img = torch.empty(new_cap, d, dtype=torch.bfloat16, device=device)
        txt = torch.empty(new_cap, d, dtype=torch.bfloat16, device=device)
        if self.img is not None and self.n:
            img[:self.n].copy_(self.img[:self.n])
            txt[:self.n].copy_(self.txt[:self.n])
        self.img, self.txt = img, txt
        if self.lab is not None:
            lab = torch.empty(new_cap, dtype=self.lab.dtype, device=device)
            lab[:self.n].copy_(self.lab[:self.n])
            self.lab = lab

    def reserve(self, rows: int, d: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
        """Views of the next ``rows`` rows of the image / text buffers for the prologue to fill."""
        self._grow(self.n + rows, d, torch.device(device), None)
        self._pending = (self.n, rows)
        return self.img[self.n:self.n + rows], self.txt[self.n:self.n + rows]

    def commit(self, image_embeddings: torch.Tensor, text_embeddings: torch.Tensor, labels: torch.Tensor,
               written: bool = False) -> None:
        """Append one step.  ``written``: the rows handed out by the last ``reserve`` already hold these
        embeddings (the prologue wrote them); otherwise they are copied (detached, rounded to bf16)."""
        rows, d = image_embeddings.shape
        dev = image_embeddings.device
        if not (written and self._pending == (self.n, rows) and self.img is not None
                and self.img.shape[1] == d and self.img.device == dev):
            self._grow(self.n + rows, d, dev, labels)
            self.img[self.n:self.n + rows].copy_(image_embeddings.detach())
            self.txt[self.n:self.n + rows].copy_(text_embeddings.detach())
        if self.lab is None or self.lab.dtype != labels.dtype or self.lab.device != dev:
            old = self.lab
            self.lab = torch.empty(self.img.shape[0], dtype=labels.dtype, device=dev)
            if old is not None and self.n:
                self.lab[:self.n].copy_(old[:self.n])
        self.lab[self.n:self.n + rows].copy_(labels.detach())
        self.n += rows
        self._pending = None

    # ------------------------------------------------------------------ reading
    def __len__(self) -> int:
        return self.n

    def get(self):
        """(image embeddings [n, d] bf16, text embeddings [n, d] bf16, labels [n]): views, no copy."""
        if self.n == 0 or self.img is None or self.lab is None:
            raise ValueError("No cached embeddings and labels")
        return self.img[:self.n], self.txt[:self.n], self.lab[:self.n]

    # the reference keeps a dict with these keys; mirror the read side of that interface
    def __contains__(self, key) -> bool:
        return self.n > 0 and key in ("image_embedding", "text_embedding", "label")

    def __getitem__(self, key):
        i, t, l = self.get()
        return {"image_embedding": i, "text_embedding": t, "label": l}[key]
