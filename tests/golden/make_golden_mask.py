"""Golden vector for the duplicate-caption mask (f3), produced by executing the REFERENCE'S OWN
``VisionLanguageModule._get_mask`` (lines 506-530), extracted with ``ast`` exactly like
``make_golden.py`` does for ``forward`` / ``_compute_loss``.  Output: ``tests/golden/mask_captions.npz``
(the caption strings and the mask the reference returns for them).

Run here (CPU container):  python tests/golden/make_golden_mask.py
"""
import os
import sys

import numpy as np
import torch

OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, OUT)
from make_golden import REF, extract_methods  # noqa: E402


class StubSelf:
    device = torch.device("cpu")


def main():
    methods, digest = extract_methods(REF, "VisionLanguageModule", {"_get_mask"})
    code, lo, hi = methods["_get_mask"]
    ns = {"torch": torch}
    exec(compile(code, f"{REF}:{lo}-{hi}", "exec"), ns)
    rng = np.random.default_rng(11)
    vocab = [f"radiograph of the {part} showing {what}" for part in ("femur", "tibia", "humerus", "pelvis")
             for what in ("an osteosarcoma", "an enchondroma", "no lesion", "a giant cell tumor", "an osteochondroma")]
    captions = [vocab[i] for i in rng.integers(0, len(vocab), size=96)]
    captions += ["a caption that occurs once", "another caption that occurs once"]
    mask = ns["_get_mask"](StubSelf(), captions)
    np.savez_compressed(os.path.join(OUT, "mask_captions.npz"), captions=np.array(captions), mask=mask.numpy(),
                        lines=np.array([lo, hi]), reference_sha256=np.array(digest))
    print("captions", len(captions), "masked pairs", int((mask == 0).sum()), "lines", lo, hi)


if __name__ == "__main__":
    main()
