"""emernerf_b200 -- B200-native (sm_100a) volumetric renderer for EmerNeRF's per-ray-batch hot path.

Drop-in for the reference's ``radiance_fields`` package and ``third_party.nerfacc_prop_net`` /
``third_party.tcnn_modules`` modules: same class names, constructor arguments, ``forward`` /
``render_rays`` / ``sampling`` signatures and state-dict keys, evaluated by hand-written CUDA kernels
behind the C ABI of ``include/emer_b200.h`` (``emernerf_b200/lib/libemer_b200.so``).

    import emernerf_b200
    emernerf_b200.install_dropin()          # before `train_emernerf.py` imports radiance_fields
    from radiance_fields import RadianceField, DensityField, build_density_field
    from radiance_fields.render_utils import render_rays
    from third_party.nerfacc_prop_net import PropNetEstimator, get_proposal_requires_grad_fn
"""
from __future__ import annotations

import sys

__version__ = "0.1.0"


def install_dropin() -> None:
    """Alias the reference's module names to this package in ``sys.modules`` so that
    ``train_emernerf.py`` / ``builders.py`` import the B200 path without modification."""
    import importlib
    import types

    from . import radiance_fields as rf
    from .third_party import nerfacc_prop_net, tcnn_modules

    sys.modules["radiance_fields"] = rf
    for sub in ("encodings", "mlp", "nerf_utils", "radiance_field", "render_utils"):
        sys.modules[f"radiance_fields.{sub}"] = importlib.import_module(f"{__name__}.radiance_fields.{sub}")
    tp = sys.modules.get("third_party")
    if tp is None:
        tp = types.ModuleType("third_party")
        tp.__path__ = []
        sys.modules["third_party"] = tp
    sys.modules["third_party.nerfacc_prop_net"] = nerfacc_prop_net
    sys.modules["third_party.tcnn_modules"] = tcnn_modules
    tp.nerfacc_prop_net = nerfacc_prop_net
    tp.tcnn_modules = tcnn_modules


def library_path() -> str:
    from . import _lib

    return _lib.lib_path()
