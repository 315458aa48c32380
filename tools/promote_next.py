"""Promote the staged kernels (csrc/next/) to the shipped sources (csrc/) with a chosen set of switches.

    python tools/promote_next.py VLP_X_UNROLL VLP_EPI_WARPS=8 [...]     # switches that won on the B200
    python tools/promote_next.py --dry-run VLP_BWD_PINGPONG

What it does (nothing else touches csrc/*.cu):
  * copies csrc/next/{lse_fwd.cu, grad_bwd.cu, grad_bwd_quad.cuh, pipeline_exp.cuh} to csrc/ with the
    include paths rewritten for the new location;
  * writes csrc/variant_defaults.cuh with one `#define` per chosen switch and includes it first,
    so the shipped build (no -D flags) compiles exactly the variant that was measured;
  * rebuilds csrc/libvlpclip.so.
Afterwards: `pytest -m gpu`, `bench.py`, an `ncu --set full` capture of the changed kernel, and
refresh profiles/ (DESIGN.md section 9).  `git diff` shows precisely what changed; `git checkout
-- <csrc files>` undoes it.
"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vlp_b200  # noqa: E402,F401
from vlp_b200 import _build  # noqa: E402

KNOWN = ("VLP_BWD_PINGPONG", "VLP_FWD_PINGPONG", "VLP_FWD_PAIR", "VLP_BWD_QUAD", "VLP_X_UNROLL",
         "VLP_EPI_WARPS", "VLP_G_SLOTS", "VLP_PUSH_SPLIT", "VLP_P_KB_PER_STAGE", "VLP_C_Q_PER_STAGE",
         "VLP_FWD_KB_PER_STAGE")
FILES = ("lse_fwd.cu", "grad_bwd.cu", "grad_bwd_quad.cuh", "pipeline_exp.cuh")


def main():
    args = sys.argv[1:]
    dry = "--dry-run" in args
    switches = [a for a in args if a != "--dry-run"]
    for sw in switches:
        name = sw.split("=")[0]
        if name not in KNOWN:
            raise SystemExit(f"unknown switch {name!r}; real-variant switches are: {', '.join(KNOWN)}")
        if name.startswith("VLP_EXP_"):
            raise SystemExit("timing mocks compute garbage and are never promoted")
    csrc = _build.CSRC
    nxt = os.path.join(csrc, "next")
    defaults = ["// variant_defaults.cuh -- switches of the staged kernels that were promoted to the shipped build",
                "// (written by tools/promote_next.py; measured with tools/pipeline_experiments.py)", "#pragma once"]
    for sw in switches:
        name, _, val = sw.partition("=")
        defaults.append(f"#ifndef {name}\n#define {name}{(' ' + val) if val else ''}\n#endif")
    out = {"variant_defaults.cuh": "\n".join(defaults) + "\n"}
    for fn in FILES:
        src = open(os.path.join(nxt, fn)).read()
        src = src.replace('#include "../common.cuh"', '#include "variant_defaults.cuh"\n#include "common.cuh"')
        src = src.replace('#include "../../../include/vlpclip.h"', '#include "../../include/vlpclip.h"')
        if fn == "pipeline_exp.cuh":
            src = re.sub(r"// csrc/next/ holds the NEXT versions.*?shipped file\.\n", "", src, flags=re.S)
        out[fn] = src
    for fn, text in out.items():
        path = os.path.join(csrc, fn)
        old = open(path).read() if os.path.exists(path) else None
        state = "unchanged" if old == text else ("new" if old is None else "changed")
        print(f"{'would write' if dry else 'writing'} {os.path.relpath(path, ROOT)} ({state})")
        if not dry:
            open(path, "w").write(text)
    if not dry:
        print("rebuilding", _build.build(force=True))
        print("next: pytest -m gpu; bench.py; ncu of the changed kernel; refresh profiles/")


if __name__ == "__main__":
    main()
