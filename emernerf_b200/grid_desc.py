"""Level table of a multi-resolution hash grid (host side of ``emer_grid_desc``).

Stands in for what ``_C.create_encoding(n_input_dims, encoding_config, precision)`` builds inside
tiny-cuda-nn (third_party/tcnn_modules.py:420-423 of the reference).  The arithmetic follows
tiny-cuda-nn's published level rules; it is deliberately computed on the host in fp32 so that the
kernel and any checker use bit-identical per-level scales.
"""
from __future__ import annotations

import ctypes
import math
from typing import Dict, List

import numpy as np

MAX_LEVELS = 16


class EmerGridDesc(ctypes.Structure):
    _fields_ = [
        ("n_dims", ctypes.c_int32),
        ("n_levels", ctypes.c_int32),
        ("n_feat", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
        ("scale", ctypes.c_float * MAX_LEVELS),
        ("resolution", ctypes.c_uint32 * MAX_LEVELS),
        ("offset", ctypes.c_uint32 * (MAX_LEVELS + 1)),
        ("hashed", ctypes.c_uint32 * MAX_LEVELS),
    ]


class GridDesc:
    """Python view + ctypes struct of one grid's level table."""

    def __init__(self, n_dims: int, config: Dict):
        otype = config.get("otype", "HashGrid")
        if otype != "HashGrid":
            raise NotImplementedError(f"encoding otype {otype!r} is not on the EmerNeRF hot path")
        if str(config.get("interpolation", "linear")).lower() != "linear":
            raise NotImplementedError("only linear interpolation is implemented")
        if n_dims not in (3, 4):
            raise ValueError(f"n_input_dims must be 3 or 4, got {n_dims}")
        L = int(config["n_levels"])
        F = int(config["n_features_per_level"])
        if not (1 <= L <= MAX_LEVELS):
            raise ValueError(f"n_levels must be in [1, {MAX_LEVELS}]")
        if F not in (1, 2, 4):
            raise ValueError("n_features_per_level must be 1, 2 or 4")
        log2_t = int(config["log2_hashmap_size"])
        base = np.float32(int(config["base_resolution"]))
        growth = np.float32(config.get("per_level_scale", 2.0))
        log2_growth = np.float32(np.log2(growth))

        self.n_dims, self.n_levels, self.n_feat = n_dims, L, F
        self.scales: List[float] = []
        self.resolutions: List[int] = []
        self.offsets: List[int] = [0]
        self.hashed: List[bool] = []
        cap = 1 << log2_t
        for level in range(L):
            scale = np.float32(np.float32(np.exp2(np.float32(level) * log2_growth)) * base - np.float32(1.0))
            res = int(math.ceil(float(scale))) + 1
            dense = res ** n_dims
            half_u32 = 0xFFFFFFFF // 2
            entries = half_u32 if float(res) ** n_dims > float(half_u32) else dense
            entries = min(((entries + 7) // 8) * 8, cap)
            # dense index is used while the running stride still fits the level
            stride, dims_indexed = 1, 0
            while dims_indexed < n_dims and stride <= entries:
                stride *= res
                dims_indexed += 1
            self.scales.append(float(scale))
            self.resolutions.append(res)
            self.offsets.append(self.offsets[-1] + entries)
            self.hashed.append(entries < stride)
        if self.offsets[-1] >= 2 ** 32:
            raise ValueError("grid too large for 32-bit entry offsets")

        c = EmerGridDesc()
        c.n_dims, c.n_levels, c.n_feat, c.reserved = n_dims, L, F, 0
        for i in range(L):
            c.scale[i] = self.scales[i]
            c.resolution[i] = self.resolutions[i]
            c.hashed[i] = 1 if self.hashed[i] else 0
        for i in range(L + 1):
            c.offset[i] = self.offsets[i]
        self.c = c

    @property
    def n_entries(self) -> int:
        return self.offsets[-1]

    @property
    def n_params(self) -> int:
        return self.offsets[-1] * self.n_feat

    @property
    def n_output_dims(self) -> int:
        return self.n_levels * self.n_feat

    def bytes_per_point(self) -> int:
        """Algorithmic bytes per encoded point: corners + position + output."""
        return self.n_levels * (2 ** self.n_dims) * self.n_feat * 4 + self.n_dims * 4 + self.n_output_dims * 4
