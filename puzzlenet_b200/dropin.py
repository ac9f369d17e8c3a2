"""Register puzzlenet_b200's modules under the reference's module names.

The reference binds its dependencies by *module name at import time* (SURVEY.md §8b):
``model5_b.py:39`` does ``import pointnet_util as pu`` and stores ``pu.sample_and_group`` at
construction (:427-429); ``model5_b.py:48`` does ``from PyTorchEMD.emd import earth_mover_distance``
and ``PyTorchEMD/emd.py:2`` does ``import emd_cuda``; ``train.py:20`` / ``test.py:22`` do
``import model5_b``.  Calling :func:`install` *before* those imports makes the reference's
``train.py`` / ``test.py`` pick up the B200 implementations without editing them.
"""
from __future__ import annotations

import sys
import types


def install(model: bool = True) -> None:
    """``model=False`` replaces only the operator modules (pointnet_util, PyTorchEMD.emd, emd_cuda) and
    leaves the reference's own ``model5_b`` in charge of the network."""
    from . import emd, emd_cuda, pointnet_util
    sys.modules["pointnet_util"] = pointnet_util
    sys.modules["emd_cuda"] = emd_cuda
    pkg = types.ModuleType("PyTorchEMD")
    pkg.__path__ = []
    pkg.emd = emd
    sys.modules["PyTorchEMD"] = pkg
    sys.modules["PyTorchEMD.emd"] = emd
    sys.modules["emd"] = emd              # PyTorchEMD/test_emd_loss.py:4 imports it as a top-level module
    if model:
        from . import model5_b
        sys.modules["model5_b"] = model5_b


def uninstall() -> None:
    for name in ("pointnet_util", "emd_cuda", "PyTorchEMD.emd", "emd", "model5_b"):
        mod = sys.modules.get(name)
        if mod is not None and getattr(mod, "__name__", "").startswith("puzzlenet_b200"):
            del sys.modules[name]
    pkg = sys.modules.get("PyTorchEMD")
    if pkg is not None and getattr(pkg, "__file__", None) is None:
        del sys.modules["PyTorchEMD"]
