"""Synthetic Waymo-shape ray batches (no dataset in this environment).

Mirrors what ``ScenePixelSource.get_train_rays`` (datasets/base/pixel_source.py:666-731, rays from
``get_rays`` :39-76) and the lidar source (datasets/base/lidar_source.py:281-308) hand to
``render_rays``: three pinhole cameras (front-left, front, front-right; 640x960, fx=fy~1030,
+-45 deg yaw) on an ego vehicle advancing along +x over ``num_timesteps`` frames, unit view
directions, per-ray image index / normalised timestamp / pixel coordinates (y/H, x/W), RGB and
sky-mask targets, optional 64-d feature targets; lidar rays with ranges in (0.5, 80) m.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

H, W = 640, 960
FX = FY = 1030.0
CX, CY = 480.0, 320.0
YAWS = (math.radians(45.0), 0.0, math.radians(-45.0))     # front_left, front, front_right


def _cam_dirs(cam: torch.Tensor, px: torch.Tensor, py: torch.Tensor) -> torch.Tensor:
    # camera frame: x right, y down, z forward  ->  vehicle frame: x front, y left, z up
    dx = (px + 0.5 - CX) / FX
    dy = (py + 0.5 - CY) / FY
    fwd, left, up = torch.ones_like(dx), -dx, -dy
    yaw = torch.tensor(YAWS, dtype=torch.float32)[cam]
    c, s = torch.cos(yaw), torch.sin(yaw)
    d = torch.stack([c * fwd - s * left, s * fwd + c * left, up], -1)
    return d / d.norm(dim=-1, keepdim=True)


def pixel_batch(n_rays: int = 8192, num_timesteps: int = 200, num_cams: int = 3, seed: int = 0,
                features: bool = False, device="cpu", pin: bool = False) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    t_idx = torch.randint(0, num_timesteps, (n_rays,), generator=g)
    cam = torch.randint(0, num_cams, (n_rays,), generator=g)
    px = torch.randint(0, W, (n_rays,), generator=g).float()
    py = torch.randint(0, H, (n_rays,), generator=g).float()
    t = t_idx.float() / max(num_timesteps - 1, 1)
    ego = torch.stack([t * 60.0, torch.zeros(n_rays), torch.full((n_rays,), 2.0)], -1)
    out = {
        "origins": ego,
        "viewdirs": _cam_dirs(cam % 3, px, py),
        "img_idx": t_idx * num_cams + cam,
        "normed_timestamps": t,
        "pixel_coords": torch.stack([py / H, px / W], -1),
        "pixels": torch.rand(n_rays, 3, generator=g),
        "sky_masks": (torch.rand(n_rays, generator=g) < 0.2).float(),
    }
    if features:
        out["features"] = torch.rand(n_rays, 64, generator=g)
    return _place(out, device, pin)


def lidar_batch(n_rays: int = 8192, num_timesteps: int = 200, seed: int = 0, device="cpu",
                pin: bool = False) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed + 99991)
    t_idx = torch.randint(0, num_timesteps, (n_rays,), generator=g)
    t = t_idx.float() / max(num_timesteps - 1, 1)
    az = torch.rand(n_rays, generator=g) * 2 * math.pi
    el = (torch.rand(n_rays, generator=g) - 0.85) * math.radians(20.0)
    d = torch.stack([torch.cos(el) * torch.cos(az), torch.cos(el) * torch.sin(az), torch.sin(el)], -1)
    out = {
        "lidar_origins": torch.stack([t * 60.0, torch.zeros(n_rays), torch.full((n_rays,), 2.0)], -1),
        "lidar_viewdirs": d / d.norm(dim=-1, keepdim=True),
        "lidar_ranges": torch.rand(n_rays, 1, generator=g) * 79.5 + 0.5,
        "lidar_normed_timestamps": t,
    }
    return _place(out, device, pin)


def _place(d, device, pin):
    if pin and torch.cuda.is_available():
        return {k: v.pin_memory() for k, v in d.items()}
    return {k: v.to(device) for k, v in d.items()}


def bytes_of(d: Dict[str, torch.Tensor]) -> int:
    return int(sum(v.numel() * v.element_size() for v in d.values()))
