"""Golden values for the retrieval metrics (f1), produced by executing the REFERENCE'S OWN
``precision_at_k_on_image_embeddings`` (lines 364-400) and ``recall_at_k_on_image_text_retreival``
(lines 402-439), extracted with ``ast`` like ``make_golden.py`` does.  Inputs are regenerated from the
seed by the test (``oracle.make_embeddings``); the file holds the metric values.

Run here (CPU container):  python tests/golden/make_golden_retrieval.py
"""
import os
import sys

import numpy as np
import torch

OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, OUT)
from make_golden import REF, O, extract_methods  # noqa: E402

KS_RECALL = [1, 3, 5, 10]
KS_PRECISION = [3, 5, 10, 15]


def main():
    names = {"precision_at_k_on_image_embeddings", "recall_at_k_on_image_text_retreival"}
    methods, digest = extract_methods(REF, "VisionLanguageModule", names)
    ns = {"torch": torch}
    for name, (code, lo, hi) in methods.items():
        exec(compile(code, f"{REF}:{lo}-{hi}", "exec"), ns)
    out = {"reference_sha256": np.array(digest)}
    for tag, (n, d, rho, seed) in {"a": (300, 64, 0.6, 3), "b": (129, 8, 0.9, 4)}.items():
        img, txt = O.make_embeddings(n, d, rho=rho, seed=seed)
        labels = torch.from_numpy(np.random.default_rng(seed).integers(0, 7, size=n))
        r = ns["recall_at_k_on_image_text_retreival"](None, img.float(), txt.float(), KS_RECALL)
        p = ns["precision_at_k_on_image_embeddings"](None, img.float(), labels, KS_PRECISION)
        out[f"{tag}_params"] = np.array([n, d, rho, seed], dtype=np.float64)
        out[f"{tag}_recall"] = np.array([r[k] for k in KS_RECALL], dtype=np.float64)
        out[f"{tag}_precision"] = np.array([p[k] for k in KS_PRECISION], dtype=np.float64)
        print(tag, "recall", r, "precision", p)
    out["lines"] = np.array([methods["precision_at_k_on_image_embeddings"][1], methods["precision_at_k_on_image_embeddings"][2],
                             methods["recall_at_k_on_image_text_retreival"][1], methods["recall_at_k_on_image_text_retreival"][2]])
    np.savez_compressed(os.path.join(OUT, "retrieval_metrics.npz"), **out)


if __name__ == "__main__":
    main()
