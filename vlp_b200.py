"""Importable alias for the package directory ``vision-language-pretraining-for-bone-tumor-detection_b200/``.

The directory name required by the repo layout contains hyphens and cannot be written in an
``import`` statement; ``import vlp_b200`` loads that directory as the package ``vlp_b200``
(sub-modules resolve inside it, e.g. ``vlp_b200.functional``).
"""
import importlib.util as _ilu
import os as _os
import sys as _sys

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)),
                         "vision-language-pretraining-for-bone-tumor-detection_b200")
_spec = _ilu.spec_from_file_location("vlp_b200", _os.path.join(_PKG_DIR, "__init__.py"),
                                     submodule_search_locations=[_PKG_DIR])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["vlp_b200"] = _mod
_spec.loader.exec_module(_mod)
