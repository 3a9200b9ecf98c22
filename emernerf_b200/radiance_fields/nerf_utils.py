"""Scene contraction and the truncated-exp density activation with the reference's names
(radiance_fields/nerf_utils.py), as single fused kernels."""
from __future__ import annotations

from typing import Union

import torch
from torch import Tensor

from .. import _ops


def contract(x: Tensor, aabb: Tensor, ord: Union[str, int, float] = None) -> Tensor:
    """MeRF-style contraction of [-inf, inf]^3 to [0, 1]^3 (nerf_utils.py:13-28).  Only the
    inf-norm variant exists on the path (radiance_field.py:290,830)."""
    if ord not in (float("inf"), "inf"):
        raise NotImplementedError("only ord=inf is used by EmerNeRF and implemented here")
    shape = x.shape
    y = _ops.contract_raw(x.reshape(-1, 3), aabb)
    return y.view(shape)


def trunc_exp(x: Tensor) -> Tensor:
    """exp with a clamped backward (nerf_utils.py:59-75)."""
    return _ops.density_activation(x + 1.0)


def find_topk_nearby_timesteps(original: Tensor, query: Tensor, topk: int = 2, return_indices: bool = False):
    """nerf_utils.py:31-56 (only reached by the disabled temporal interpolation)."""
    diffs = (original.unsqueeze(0) - query.unsqueeze(1)).abs()
    _, idx = torch.topk(diffs, k=topk, dim=1, largest=False)
    return (original[idx], idx) if return_indices else original[idx]
