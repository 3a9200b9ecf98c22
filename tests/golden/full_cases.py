"""The BENCHMARKED configurations (BASELINE.json configs[1..4]: 2^20 / 2^18-entry tables, 8x1 proposal
grids, 64 samples, proposal samples [128, 64]) as parity cases on a 256-ray batch.

The miniature cases of ``cases.py`` run 768 rows per layer, below ``_ops.TC_MIN_ROWS``: they exercise the
CUDA-core layers and ``prop_level_kernel<4>``.  These cases run 16 384 rows through every head and 32 768 /
16 384 samples through ``prop_level_kernel<8>``, i.e. the kernels ``bench.py`` times.

The tables (122 MB for the static grid) cannot live in a fixture, so they come from a counter-based integer
hash evaluated in numpy (portable: no dependence on torch's RNG streams); every other parameter (MLPs,
embedding, PE map) is stored in the fixture.
"""
from __future__ import annotations

import types
from typing import Dict

import numpy as np
import torch

import cases

AABB = cases.AABB
N_TIMESTEPS = 200
N_CAMS = 3
N_RAYS = 256
NUM_SAMPLES = 64
PROP_SAMPLES = [128, 64]
NEAR, FAR = 0.1, 1000.0
VARIANTS = list(cases.CASES)            # static, dynamic, flow, flow_feat

# configs/default_config.yaml:62-77, radiance_field.py:916-923 (flow grid), builders.py:98-110 (proposal grids)
ENC_STATIC = dict(n_input_dims=3, n_levels=10, base_resolution=16, max_resolution=8192,
                  log2_hashmap_size=20, n_features_per_level=4)
ENC_DYN = dict(n_input_dims=4, n_levels=10, base_resolution=32, max_resolution=8192,
               log2_hashmap_size=18, n_features_per_level=4)
ENC_FLOW = dict(n_input_dims=4, n_levels=10, base_resolution=16, max_resolution=4096,
                log2_hashmap_size=18, n_features_per_level=4)
ENC_PROP = [dict(n_levels=8, max_resolution=512, log2_hashmap_size=20, n_features_per_level=1),
            dict(n_levels=8, max_resolution=2048, log2_hashmap_size=20, n_features_per_level=1)]
TABLE_AMPLITUDE = 0.5                   # uniform(-a, a): std 0.29, the "trained-like" state of bench.py


def render_cfg():
    ns = types.SimpleNamespace
    return ns(nerf=ns(sampling=ns(num_samples=NUM_SAMPLES),
                      propnet=ns(num_samples_per_prop=list(PROP_SAMPLES), near_plane=NEAR, far_plane=FAR,
                                 sampling_type="uniform_lindisp")),
              render=ns(render_chunk_size=16384))


def hashed_table(n: int, stream: int) -> torch.Tensor:
    """n floats in (-TABLE_AMPLITUDE, TABLE_AMPLITUDE) from a 64-bit finaliser of (index, stream)."""
    with np.errstate(over="ignore"):
        x = np.arange(n, dtype=np.uint64) + np.uint64(stream) * np.uint64(0x9E3779B97F4A7C15)
        x ^= x >> np.uint64(33)
        x *= np.uint64(0xFF51AFD7ED558CCD)
        x ^= x >> np.uint64(33)
        x *= np.uint64(0xC4CEB9FE1A85EC53)
        x ^= x >> np.uint64(33)
    u = (x >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / (1 << 24))        # [0, 1), 24 bits: exact
    return torch.from_numpy((u * np.float32(2.0) - np.float32(1.0)) * np.float32(TABLE_AMPLITUDE))


def build_models(ns, variant: str, seed: int = 0):
    """(field, proposal networks) of the full-size configuration from any namespace with the reference's class
    names -- the reference itself (fixture generation) or the drop-in (tests)."""
    c = cases.CASES[variant]
    torch.manual_seed(seed)
    enc = ns.HashEncoder(verbose=False, **ENC_STATIC)
    dyn = ns.HashEncoder(verbose=False, **ENC_DYN) if c["dynamic"] else None
    flw = ns.HashEncoder(verbose=False, **ENC_FLOW) if c["flow"] else None
    field = ns.RadianceField(
        xyz_encoder=enc, dynamic_xyz_encoder=dyn, flow_xyz_encoder=flw, aabb=AABB, unbounded=True,
        geometry_feature_dim=64, base_mlp_layer_width=64, head_mlp_layer_width=64,
        enable_cam_embedding=False, enable_img_embedding=True, num_cams=N_CAMS,
        appearance_embedding_dim=16, semantic_feature_dim=64, feature_mlp_layer_width=64,
        feature_embedding_dim=64, enable_sky_head=True, enable_shadow_head=c["shadow"],
        enable_feature_head=c["feature"], num_train_timesteps=N_TIMESTEPS,
        interpolate_xyz_encoding=True, enable_learnable_pe=True, enable_temporal_interpolation=False)
    field.register_normalized_training_timesteps(torch.linspace(0, 1, N_TIMESTEPS), time_diff=1.0 / N_TIMESTEPS)
    props = []
    for e in ENC_PROP:
        p = ns.build_density_field(n_input_dims=3, n_levels=e["n_levels"], max_resolution=e["max_resolution"],
                                   log2_hashmap_size=e["log2_hashmap_size"],
                                   n_features_per_level=e["n_features_per_level"], unbounded=True)
        p.set_aabb(AABB)
        props.append(p)
    fill_tables(field, props)
    return field, props


def fill_tables(field, props) -> None:
    stream = 1
    with torch.no_grad():
        for m in [field] + list(props):
            for k, v in m.named_parameters():
                if k.endswith("tcnn_encoding.params"):
                    v.copy_(hashed_table(v.numel(), stream).to(v.device))
                    stream += 1


def small_state(m) -> Dict[str, torch.Tensor]:
    """Everything of a module's state-dict except the hash tables."""
    return {k: v for k, v in m.state_dict().items() if not k.endswith("tcnn_encoding.params")}


def make_batch(variant: str, seed: int = 0, lidar: bool = False) -> Dict[str, torch.Tensor]:
    """Waymo-shape rays (emernerf_b200/synthetic.py) -- the distribution the benchmark renders."""
    from emernerf_b200 import synthetic

    if lidar:
        return synthetic.lidar_batch(N_RAYS, N_TIMESTEPS, seed=seed)
    return synthetic.pixel_batch(N_RAYS, N_TIMESTEPS, N_CAMS, seed=seed, features=cases.CASES[variant]["feature"])


def projections(grad: torch.Tensor, n_proj: int = 16, seed: int = 5) -> torch.Tensor:
    """Digest of a (huge, sparse) table gradient: its dot products with n_proj fixed +-1 vectors from the same
    portable hash, plus its L1 and L2 norms.  [n_proj + 2] float64."""
    g = grad.detach().double().cpu().reshape(-1)
    out = []
    for j in range(n_proj):
        sign = torch.sign(hashed_table(g.numel(), 1000 + seed * 64 + j).double())
        out.append((g * sign).sum())
    out += [g.abs().sum(), g.square().sum().sqrt()]
    return torch.stack(out)


def mask_rays(out: dict, keep: torch.Tensor) -> dict:
    """render_rays results restricted to the rays ``keep`` ([R] bool), extras included."""
    res = {}
    for k, v in out.items():
        if isinstance(v, dict):
            res[k] = mask_rays(v, keep)
        else:
            res[k] = v[keep.to(v.device)]
    return res


def mask_prop_cache(cache: list, keep: torch.Tensor) -> None:
    """Restrict the estimator's cached (intervals, cdfs, level) entries to the rays ``keep``, in place -- the
    proposal loss is then the reference's own ``compute_loss`` over those rays."""
    for i, (iv, cdfs, lvl) in enumerate(cache):
        iv.vals = iv.vals[keep.to(iv.vals.device)]
        cache[i] = (iv, None if cdfs is None else cdfs[keep.to(cdfs.device)], lvl)
