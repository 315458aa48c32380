"""Generate golden input/output vectors by executing the REFERENCE'S OWN CODE.

The reference module cannot be imported in this image (lightning / timm / monai / torchmetrics /
hydra are not installed) and it ships no tests or fixtures for the CLIP head.  This script pins the
oracle anyway: it parses ``/root/reference/src/models/pretrain/VisionLanguageModule.py`` with
``ast``, extracts the source of ``VisionLanguageModule.forward`` (lines 441-461) and
``VisionLanguageModule._compute_loss`` (lines 532-554), compiles exactly that source and runs it
bound to a stub ``self`` whose encoders are identities.  Autograd through the reference code yields
the reference gradients.  Outputs are written to ``tests/golden/*.npz`` (committed; they travel to
the GPU box, ``/root/reference`` does not).

Run here (CPU container):  python tests/golden/make_golden.py
"""
import ast
import hashlib
import math
import os
import sys
import textwrap
import types

import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference/src/models/pretrain/VisionLanguageModule.py"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
from oracle import clip_oracle as O  # noqa: E402  (only for the seeded input generators)


def extract_methods(path, cls_name, names):
    src = open(path).read()
    tree = ast.parse(src)
    out = {}
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == cls_name:
            for item in node.body:
                if isinstance(item, ast.FunctionDef) and item.name in names:
                    seg = ast.get_source_segment(src, item)
                    out[item.name] = (textwrap.dedent(seg), item.lineno, item.end_lineno)
    return out, hashlib.sha256(src.encode()).hexdigest()


def build_reference_callables():
    methods, digest = extract_methods(REF, "VisionLanguageModule", {"forward", "_compute_loss"})
    ns = {"torch": torch, "F": F, "np": np}
    for name, (code, lo, hi) in methods.items():
        exec(compile(code, f"{REF}:{lo}-{hi}", "exec"), ns)
    return ns["forward"], ns["_compute_loss"], methods, digest


class StubSelf:
    """`self` for the extracted methods: identity encoders + the three head parameters."""

    def __init__(self, w_img, w_txt, logit_scale):
        self.image_projection = w_img
        self.text_projection = w_txt
        self.logit_scale = logit_scale

    @staticmethod
    def image_encoder(x):
        return x

    @staticmethod
    def text_encoder(features=None):
        return features


def run_case(forward, compute_loss, n, f_img, f_txt, d, logit_scale, seed, dtype):
    fi, ft, wi, wt = O.make_features(n, f_img, f_txt, d, seed=seed)
    fi, ft, wi, wt = (t.to(dtype).requires_grad_(True) for t in (fi, ft, wi, wt))
    ls = torch.tensor([logit_scale], dtype=torch.float64, requires_grad=True)  # reference :111 is fp64
    stub = StubSelf(wi, wt, ls)
    batch = {"x-ray": fi, "caption_tokenized": {"features": ft}}
    logits, ie, te = forward(stub, batch)
    loss, il, tl = compute_loss(stub, logits, False, False, None)
    loss.backward()
    return {"image_features": fi, "text_features": ft, "image_projection": wi, "text_projection": wt,
            "logit_scale": ls, "logits": logits, "image_embeddings": ie, "text_embeddings": te,
            "loss": loss, "image_loss": il, "text_loss": tl,
            "d_image_features": fi.grad, "d_text_features": ft.grad,
            "d_image_projection": wi.grad, "d_text_projection": wt.grad, "d_logit_scale": ls.grad}


def run_embedding_case(compute_loss, n, d, rho, logit_scale, seed):
    """Embedding-level parity surface: bf16-representable unit embeddings, fp32 like the spec."""
    I, T = O.make_embeddings(n, d, rho=rho, seed=seed)
    I = I.clone().requires_grad_(True)
    T = T.clone().requires_grad_(True)
    ls = torch.tensor([logit_scale], dtype=torch.float64, requires_grad=True)
    scale = torch.clamp(ls.exp(), max=100)              # reference :456-457
    logits = (I @ T.T) * scale                          # reference :459
    loss, il, tl = compute_loss(None, logits, False, False, None)
    loss.backward()
    return {"I": I, "T": T, "logit_scale": ls, "loss": loss, "image_loss": il, "text_loss": tl,
            "dI": I.grad, "dT": T.grad, "d_logit_scale": ls.grad}


def to_np(d):
    return {k: v.detach().cpu().numpy() for k, v in d.items()}


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    forward, compute_loss, methods, digest = build_reference_callables()
    meta = {"reference_sha256": digest,
            "forward_lines": list(methods["forward"][1:]),
            "compute_loss_lines": list(methods["_compute_loss"][1:]),
            "torch": torch.__version__}
    print(meta)
    head_cases = {
        "head_n32_f512_312_d128_fp32": dict(n=32, f_img=512, f_txt=312, d=128, logit_scale=math.log(1 / 0.07), seed=42, dtype=torch.float32),
        "head_n64_f128_40_d64_fp64": dict(n=64, f_img=128, f_txt=40, d=64, logit_scale=math.log(1 / 0.07), seed=43, dtype=torch.float64),
        "head_n48_f64_40_d32_clamped": dict(n=48, f_img=64, f_txt=40, d=32, logit_scale=5.0, seed=44, dtype=torch.float64),
    }
    for name, kw in head_cases.items():
        res = to_np(run_case(forward, compute_loss, **kw))
        # inputs are regenerated from the seed by the tests (oracle.make_features); keep checksums
        for key in ("image_features", "text_features", "image_projection", "text_projection"):
            v = res.pop(key).astype(np.float64)
            res[key + "_checksum"] = np.array([v.sum(), np.abs(v).sum()])
        res["params"] = np.array([kw["n"], kw["f_img"], kw["f_txt"], kw["d"], kw["logit_scale"], kw["seed"],
                                  64 if kw["dtype"] == torch.float64 else 32], dtype=np.float64)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **res)
        print(name, "loss", float(res["loss"]))
    emb_cases = {
        "emb_n256_d512_init_rho035": dict(n=256, d=512, rho=0.35, logit_scale=math.log(1 / 0.07), seed=42),
        "emb_n128_d256_init_rho0": dict(n=128, d=256, rho=0.0, logit_scale=math.log(1 / 0.07), seed=42),
        "emb_n256_d512_ln50_rho035": dict(n=256, d=512, rho=0.35, logit_scale=math.log(50.0), seed=42),
        "emb_n128_d256_clamped_rho0": dict(n=128, d=256, rho=0.0, logit_scale=5.0, seed=42),
        "emb_n100_d72_ragged": dict(n=100, d=72, rho=0.35, logit_scale=3.0, seed=5),
    }
    for name, kw in emb_cases.items():
        res = to_np(run_embedding_case(compute_loss, **kw))
        # inputs are regenerated from the seed by the tests; keep the file small
        small = {k: v for k, v in res.items() if k not in ("I", "T")}
        small["I_checksum"] = np.array([res["I"].astype(np.float64).sum(), np.abs(res["I"]).astype(np.float64).sum()])
        small["T_checksum"] = np.array([res["T"].astype(np.float64).sum(), np.abs(res["T"]).astype(np.float64).sum()])
        small["params"] = np.array([kw["n"], kw["d"], kw["rho"], kw["logit_scale"], kw["seed"]], dtype=np.float64)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **small)
        print(name, "loss", float(res["loss"]))
    # deprecated flags must raise exactly like the reference (lines 535-547)
    for flags in ((True, False), (False, True)):
        try:
            compute_loss(None, torch.zeros(2, 2), flags[0], flags[1], ["a", "b"])
            raise SystemExit("reference did not raise for deprecated flags")
        except DeprecationWarning as exc:
            meta[f"deprecation_{int(flags[0])}{int(flags[1])}"] = str(exc)
    import json
    with open(os.path.join(OUT, "golden_meta.json"), "w") as fh:
        json.dump(meta, fh, indent=1)


if __name__ == "__main__":
    main()
