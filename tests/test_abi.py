"""CPU: the C-ABI shared library builds for sm_100a, loads without a GPU and exports every symbol
``include/vlpclip.h`` declares (no compute calls here)."""
import ctypes
import os
import re
import subprocess

import pytest

import vlp_b200  # noqa: F401
from vlp_b200 import _build, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vlpclip.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vlpclip_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    fns = header_functions()
    for required in ("vlpclip_lse_fwd", "vlpclip_lse_merge", "vlpclip_loss_reduce", "vlpclip_grad",
                     "vlpclip_project_normalize_fwd", "vlpclip_normalize_bwd", "vlpclip_gemm_tf32",
                     "vlpclip_last_error", "vlpclip_version"):
        assert required in fns


def test_library_builds_and_exports_every_declared_symbol():
    path = _build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    missing = [f for f in header_functions() if not hasattr(lib, f)]
    assert not missing, f"declared in vlpclip.h but not exported: {missing}"


def test_ctypes_signatures_cover_the_header():
    assert sorted(_lib.declared_symbols()) == header_functions()


def test_version_and_error_string_without_gpu():
    lib = _lib.load()
    assert lib.vlpclip_version() == 100
    # size queries are pure host arithmetic
    assert lib.vlpclip_lse_workspace_bytes(256, 256, 512) > 0
    assert lib.vlpclip_grad_workspace_bytes(256, 256, 512) > 0
    assert lib.vlpclip_lse_workspace_bytes(0, 256, 512) == 0
    # argument validation happens before any CUDA call and reports through last_error
    rc = lib.vlpclip_lse_merge(None, None, None, 0, 0, None, None, None, None, None, None, None, None)
    assert rc == -1 and b"empty" in lib.vlpclip_last_error()
    with pytest.raises(ValueError):
        _lib.check(rc, "lse_merge")


def test_kernels_are_blackwell_native():
    """SASS must contain tcgen05 MMA / TMEM / TMA instructions and no legacy HMMA path."""
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _build.build()], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass       # tcgen05.mma kind::f16
    assert "LDTM" in sass and "STTM" in sass   # tcgen05.ld / st
    assert "UTMALDG" in sass       # cp.async.bulk.tensor
    assert "UBLKCP" in sass        # cp.async.bulk (DSMEM tile push)
    assert "HMMA." not in sass.replace("UTCHMMA", "")


# ---- backward work partition (host-side view of the kernel's "stream-K" split) ---------------
def _grad_plan(n_row_blocks, tiles, n_clusters):
    import ctypes

    import numpy as np
    lib = _lib.load()
    cap = n_row_blocks * tiles + 4 * n_clusters
    seg = np.zeros((cap, 5), np.int32)
    red = np.zeros((cap, 3), np.int32)
    ns, nr = ctypes.c_int(), ctypes.c_int()
    rc = lib.vlpclip_grad_plan(n_row_blocks, tiles, n_clusters, seg.ctypes.data, cap,
                               ctypes.byref(ns), red.ctypes.data, cap, ctypes.byref(nr))
    assert rc == 0, lib.vlpclip_last_error()
    return seg[:ns.value], red[:nr.value]


@pytest.mark.parametrize("shape", [(256, 256, 74), (2, 2, 74), (1, 1, 1), (32, 256, 72),
                                   (256, 32, 72), (3, 5, 4), (512, 512, 74), (7, 1, 3),
                                   (1, 300, 74), (128, 256, 72), (5, 7, 74),
                                   # the staged quad kernel cuts (row-block PAIRS x tiles) over 4-CTA clusters
                                   (128, 256, 37), (16, 256, 37), (128, 32, 37), (4, 8, 37), (128, 256, 33)])
def test_backward_work_partition_is_exact_and_balanced(shape):
    """Every (row block, column tile) is swept exactly once, the load differs by <= 1 tile between
    SM pairs, each pair owns at most one head and one tail partial block, and the pieces the
    reduce kernel sums for a split row block are exactly the partial segments the sweep wrote."""
    import numpy as np
    r, c, p = shape
    seg, red = _grad_plan(r, c, p)
    cover = np.zeros((r, c), int)
    load = {}
    for cl, rb, t0, t1, slot in seg:
        assert 0 <= t0 < t1 <= c
        cover[rb, t0:t1] += 1
        load[cl] = load.get(cl, 0) + (t1 - t0)
        assert (slot == -1) == (t0 == 0 and t1 == c)
    assert (cover == 1).all()
    assert max(load.values()) - min(load.values()) <= 1
    partial = [(rb, cl, slot) for cl, rb, t0, t1, slot in seg if slot >= 0]
    assert len({(cl, slot) for _, cl, slot in partial}) == len(partial)     # slots never reused
    assert sorted(partial) == sorted(map(tuple, red.tolist()))
    # summation order inside a row block = ascending cluster index (fixed => reproducible)
    for rb in set(red[:, 0].tolist()):
        cls = red[red[:, 0] == rb][:, 1]
        assert (np.diff(cls) > 0).all()
