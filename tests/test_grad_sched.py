"""CPU: the schedule of the single-recompute backward (csrc/grad_sched.cuh, vlpclip_grad_both_plan).

Two layers: (1) tests/sched_check.cpp, a brute-force C++ check compiled with g++ against the very
header the kernel includes (exact cover, one tile per producer and step, distinct columns per step,
monotone consumer time, piece ranks, and an event simulation of the ring protocol that must not
deadlock); (2) the host-side plan exported by the library, checked from Python for the shapes the
bench and the sharded job use."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "vision-language-pretraining-for-bone-tumor-detection_b200", "csrc")


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_schedule_brute_force(tmp_path):
    exe = str(tmp_path / "sched_check")
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-I", CSRC, os.path.join(ROOT, "tests", "sched_check.cpp"),
                        "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "SCHED CHECK PASSED" in r.stdout


def _plan(lib, R, C, np_, nq):
    n = R * C
    info = (ctypes.c_int * 8)()
    prod = np.zeros((n, 5), dtype=np.int32)
    cons = np.zeros((n, 7), dtype=np.int32)
    n_prod, n_cons = ctypes.c_int(), ctypes.c_int()
    rc = lib.vlpclip_grad_both_plan(R, C, np_, nq, info, prod.ctypes.data, n, ctypes.byref(n_prod),
                                    cons.ctypes.data, n, ctypes.byref(n_cons))
    assert rc == 0, lib.vlpclip_last_error()
    assert n_prod.value == n and n_cons.value == n
    return list(info), prod, cons


@pytest.mark.parametrize("R,C", [(256, 256), (32, 256), (64, 256), (128, 256), (2, 2), (51, 8), (512, 512)])
def test_plan_exported_by_the_library(R, C):
    import vlp_b200  # noqa: F401
    from vlp_b200 import _lib
    lib = _lib.load()
    info, prod, cons = _plan(lib, R, C, 49, 50)
    n_ph, steps, n_parts = info[0], info[1], info[2]
    assert 1 <= n_ph <= 2
    # producers: exact cover of the tile grid, one tile per (slot, step), distinct columns per step
    assert len({(rb, col) for _, _, rb, col, _ in prod.tolist()}) == R * C
    assert len({(a, t) for a, t, _, _, _ in prod.tolist()}) == R * C
    assert len({(t, col) for _, t, _, col, _ in prod.tolist()}) == R * C
    assert prod[:, 1].max() < steps
    parts = {p for p in prod[:, 4].tolist() if p >= 0}
    assert parts == set(range(n_parts))
    # consumers see exactly the producers' tiles, in increasing nominal time per consumer
    assert {(a, t, rb, col) for _, a, t, rb, col, _, _ in cons.tolist()} == \
           {(a, t, rb, col) for a, t, rb, col, _ in prod.tolist()}
    for q in range(50):
        tq = cons[cons[:, 0] == q][:, 2]
        assert (np.diff(tq) > 0).all()
    # pieces of a column: ranks 0..total-1, every rank's tiles later than the previous rank's
    for col in range(C):
        cc = cons[cons[:, 4] == col]
        total = cc[0, 6]
        assert (cc[:, 6] == total).all() and set(cc[:, 5].tolist()) == set(range(total))
        last = -1
        for k in range(total):
            tk = cc[cc[:, 5] == k][:, 2]
            assert tk.min() > last
            last = tk.max()
    # the headline shape keeps the producer slots busy
    if (R, C) == (256, 256):
        assert steps <= 1.01 * R * C / 49
