"""Shared test helpers: golden loading and tensor comparison."""
from __future__ import annotations

import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Golden:
    def __init__(self, case: str):
        self.case = case
        self.z = np.load(os.path.join(GOLDEN_DIR, f"{case}.npz"))

    def tensors(self, prefix: str, device="cpu"):
        out = {}
        pre = prefix.rstrip("/") + "/"
        for k in self.z.files:
            if k.startswith(pre):
                out[k[len(pre):]] = torch.from_numpy(self.z[k]).to(device)
        return out

    def nested(self, prefix: str, device="cpu"):
        """'train/out' -> {'rgb':..., 'extras': {...}}"""
        flat = self.tensors(prefix, device)
        out = {}
        for k, v in flat.items():
            if "/" in k:
                a, b = k.split("/", 1)
                out.setdefault(a, {})[b] = v
            else:
                out[k] = v
        return out

    def scalar(self, key):
        return float(self.z[key])

    def jitters(self, mode, device="cpu"):
        js, i = [], 0
        while f"{mode}/jitter{i}" in self.z.files:
            js.append(torch.from_numpy(self.z[f"{mode}/jitter{i}"]).to(device))
            i += 1
        return js or None

    def noise(self, mode, device="cpu"):
        k = f"{mode}/noise"
        return torch.from_numpy(self.z[k]).to(device) if k in self.z.files else None


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def assert_close_dict(got, want, tol, path="", skip=()):
    assert set(got) == set(want), f"{path}: keys differ {set(got) ^ set(want)}"
    for k in want:
        if k in skip:
            continue
        if isinstance(want[k], dict):
            assert_close_dict(got[k], want[k], tol, path + k + "/", skip)
            continue
        assert tuple(got[k].shape) == tuple(want[k].shape), f"{path}{k}: {got[k].shape} vs {want[k].shape}"
        t = tol[k] if isinstance(tol, dict) and k in tol else (tol["*"] if isinstance(tol, dict) else tol)
        e = rel_err(got[k], want[k])
        assert e <= t, f"{path}{k}: rel err {e:.3e} > {t:.1e}"
