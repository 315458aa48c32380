"""ctypes binding of ``csrc/libvlpclip.so`` (the C ABI declared in ``include/vlpclip.h``).

The library is the only compute path of this package: if it cannot be loaded the import of any
op raises -- there is no CPU or eager-PyTorch fallback.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_size_t, c_void_p

from . import _build

_lib = None

_SIGNATURES = {
    "vlpclip_version": (c_int, []),
    "vlpclip_last_error": (c_char_p, []),
    "vlpclip_sm_count": (c_int, []),
    "vlpclip_launch_count": (ctypes.c_ulonglong, []),
    "vlpclip_set_sm_limit": (c_int, [c_int]),
    "vlpclip_cast_bf16_to_f16": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "vlpclip_cast_f32_operands": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vlpclip_lse_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "vlpclip_lse_fwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vlpclip_lse_fused_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "vlpclip_lse_fwd_fused": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                      c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_size_t, c_void_p]),
    "vlpclip_lse_fwd_f16": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                    c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vlpclip_lse_fwd_fused_f16": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int,
                                          c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_void_p, c_size_t, c_void_p]),
    "vlpclip_cast_push_f16": (c_int, [c_void_p, c_size_t, c_void_p, c_int, c_void_p]),
    "vlpclip_lse_merge": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "vlpclip_loss_reduce": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "vlpclip_scale_prep": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "vlpclip_loss_finish": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "vlpclip_grad_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "vlpclip_grad": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                             c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int,
                             c_int, c_float, c_float, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                             c_size_t, c_void_p]),
    "vlpclip_time_grad_kernel": (c_int, [c_int]),
    "vlpclip_last_grad_kernel_ms": (c_float, []),
    "vlpclip_dev_set_wait_profile": (c_int, [c_void_p]),
    "vlpclip_grad_plan": (c_int, [c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int,
                                  c_void_p]),
    "vlpclip_grad_scatter": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                                     c_int, c_int, c_float, c_float, c_void_p, c_int, c_int,
                                     c_void_p, c_void_p, c_size_t, c_void_p]),
    "vlpclip_grad_both_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "vlpclip_grad_both": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int,
                                  c_int, c_float, c_float, c_void_p, c_int, c_void_p, c_int, c_void_p,
                                  c_void_p, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vlpclip_grad_both_masked": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int,
                                         c_int, c_float, c_float, c_void_p, c_int, c_void_p, c_int, c_void_p,
                                         c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_size_t, c_void_p]),
    "vlpclip_lse_fwd_fused_masked": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                             c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                             c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vlpclip_grad_both_plan": (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                                       c_void_p, c_void_p, c_int, c_void_p]),
    "vlpclip_retrieval_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "vlpclip_retrieval_ranks": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                        c_void_p, c_size_t, c_void_p]),
    "vlpclip_retrieval_topk": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                       c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vlpclip_slot_sum": (c_int, [c_void_p, c_int, c_size_t, c_void_p, c_int, c_void_p, c_void_p]),
    "vlpclip_peer_alloc": (c_int, [c_size_t, c_void_p, c_void_p]),
    "vlpclip_peer_open": (c_int, [c_void_p, c_void_p]),
    "vlpclip_peer_close": (c_int, [c_void_p]),
    "vlpclip_peer_free": (c_int, [c_void_p]),
    "vlpclip_project_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "vlpclip_project_normalize_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                                              c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                              c_void_p]),
    "vlpclip_normalize_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p,
                                      c_void_p]),
    "vlpclip_gemm_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "vlpclip_gemm_tf32": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                  c_void_p, c_size_t, c_void_p]),
}


def declared_symbols():
    """Names declared in include/vlpclip.h (kept in sync by tests/test_abi.py)."""
    return sorted(_SIGNATURES)


def lib_path() -> str:
    return _build.LIB_PATH


def load(build_if_missing: bool = True) -> ctypes.CDLL:
    """Load (building first when the .so is absent/stale and nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if build_if_missing and _build.needs_build():
        try:
            _build.build()
        except Exception as exc:  # stale-but-present library is still usable on a box without nvcc
            if not os.path.exists(path):
                raise RuntimeError(
                    f"vlpclip: CUDA extension missing and could not be built ({exc}); "
                    "this package has no fallback path") from exc
    if not os.path.exists(path):
        raise RuntimeError(f"vlpclip: {path} not found; run `python __graft_entry__.py build`")
    lib = ctypes.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            continue  # using it later raises AttributeError; tests/test_abi.py checks the full set
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().vlpclip_last_error()
        msg = msg.decode() if msg else "unknown error"
        if rc == -1:
            raise ValueError(f"vlpclip.{what}: {msg}")
        raise RuntimeError(f"vlpclip.{what} failed ({rc}): {msg}")
