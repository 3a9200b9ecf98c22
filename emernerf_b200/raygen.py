"""Ray generation on the device: the producer side of ``render_rays``' ``data_dict``.

``get_rays`` keeps the reference's signature and outputs (datasets/base/pixel_source.py:39-76); ``train_rays`` is the
part of ``ScenePixelSource.get_train_rays`` (:666-731) that turns sampled (image, y, x) indices into the ray entries of
the batch -- origins, unit view directions, direction norms, pixel coordinates (y/H, x/W), normalised timestamps and
the image index -- in ONE launch (``emer_gen_rays``, csrc/raygen.cu) straight from the per-image camera tables, without
the gathered [R, 4, 4] / [R, 3, 3] copies.  A training batch then needs 3 integers per ray from the host instead of
15 floats.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from . import _lib, _ops


def _launch(img_idx, x, y, c2w, intrinsic, per_ray, timestamps, height, width, want_coords):
    _ops._need_cuda(x, y, c2w, intrinsic)
    n = x.shape[0]
    f32 = dict(dtype=torch.float32, device=x.device)
    x, y = _ops._f32c(x), _ops._f32c(y)
    c2w, intrinsic = _ops._f32c(c2w).reshape(-1, 16), _ops._f32c(intrinsic).reshape(-1, 9)
    origins, viewdirs, norms = torch.empty((n, 3), **f32), torch.empty((n, 3), **f32), torch.empty((n, 1), **f32)
    coords = torch.empty((n, 2), **f32) if want_coords else None
    ts = None if timestamps is None else _ops._f32c(timestamps)
    times = torch.empty(n, **f32) if ts is not None else None
    idx = None if img_idx is None else img_idx.to(torch.int64).contiguous()
    p = _ops._ptr
    _lib.call("emer_gen_rays", p(idx), p(x), p(y), p(c2w), p(intrinsic), int(per_ray), p(ts), int(height), int(width),
              p(origins), p(viewdirs), p(norms), p(coords), p(times), n, _ops._stream())
    return origins, viewdirs, norms, coords, times


@torch.no_grad()
def get_rays(x: Tensor, y: Tensor, c2w: Tensor, intrinsic: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """(origins [R,3], viewdirs [R,3], direction_norm [R,1]) for pixel centres (x, y); ``c2w`` [R or 1, 4, 4] (or
    [4,4]), ``intrinsic`` [R or 1, 3, 3] (or [3,3]) -- pixel_source.py:39-76."""
    c2w = c2w if c2w.dim() == 3 else c2w[None]
    intrinsic = intrinsic if intrinsic.dim() == 3 else intrinsic[None]
    n = x.shape[0]
    if c2w.shape[0] != intrinsic.shape[0]:
        c2w, intrinsic = c2w.expand(n, 4, 4), intrinsic.expand(n, 3, 3)
    if c2w.shape[0] not in (1, n):
        raise ValueError(f"get_rays: {c2w.shape[0]} camera matrices for {n} rays")
    o, d, nrm, _, _ = _launch(None, x, y, c2w, intrinsic, c2w.shape[0] == n and n > 1, None, 0, 0, False)
    return o, d, nrm


@torch.no_grad()
def train_rays(img_idx: Tensor, y: Tensor, x: Tensor, cam_to_worlds: Tensor, intrinsics: Tensor, height: int, width: int,
               normalized_timestamps: Optional[Tensor] = None) -> Dict[str, Tensor]:
    """The ray entries of ``get_train_rays``' batch from sampled indices and the per-IMAGE camera tables
    (``cam_to_worlds`` [M,4,4], ``intrinsics`` [M,3,3], ``normalized_timestamps`` [M])."""
    o, d, nrm, coords, times = _launch(img_idx, x.float(), y.float(), cam_to_worlds, intrinsics, False,
                                       normalized_timestamps, height, width, True)
    out = {"origins": o, "viewdirs": d, "direction_norms": nrm, "pixel_coords": coords,
           "img_idx": img_idx.to(torch.int64)}
    if times is not None:
        out["normed_timestamps"] = times
    return out
