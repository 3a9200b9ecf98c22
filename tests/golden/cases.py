"""Small parity cases shared by the golden generator and the tests.

Each case is a miniature of one BASELINE.json config (same module graph, tiny tables and
batches) so the CPU oracle finishes in seconds and the fixture stays small.
"""
from __future__ import annotations

import types
from typing import Dict

import torch

AABB = [-20.0, -40.0, 0.0, 80.0, 40.0, 20.0]       # configs/default_config.yaml:42
N_TIMESTEPS = 10
N_CAMS = 3

# name -> (model kwargs deltas, has lidar pass)
CASES = {
    # BASELINE configs[1] in miniature: static field, sky head, per-image embedding
    "static": dict(dynamic=False, flow=False, shadow=False, feature=False),
    # configs[2]: static + dynamic + shadow
    "dynamic": dict(dynamic=True, flow=False, shadow=True, feature=False),
    # configs[3]: + flow heads / temporal aggregation
    "flow": dict(dynamic=True, flow=True, shadow=True, feature=False),
    # configs[4]: + feature head with learnable PE
    "flow_feat": dict(dynamic=True, flow=True, shadow=True, feature=True),
}

N_RAYS = 48
NUM_SAMPLES = 16
PROP_SAMPLES = [32, 16]
NEAR, FAR = 0.1, 1000.0

ENC_STATIC = dict(n_input_dims=3, n_levels=4, base_resolution=8, max_resolution=64,
                  log2_hashmap_size=10, n_features_per_level=4)
ENC_DYN = dict(n_input_dims=4, n_levels=4, base_resolution=4, max_resolution=32,
               log2_hashmap_size=10, n_features_per_level=4)
ENC_FLOW = dict(n_input_dims=4, n_levels=4, base_resolution=4, max_resolution=24,
                log2_hashmap_size=10, n_features_per_level=4)
ENC_PROP = [dict(n_input_dims=3, n_levels=4, base_resolution=16, max_resolution=48,
                 log2_hashmap_size=12, n_features_per_level=1),
            dict(n_input_dims=3, n_levels=4, base_resolution=16, max_resolution=96,
                 log2_hashmap_size=12, n_features_per_level=1)]


def render_cfg():
    ns = types.SimpleNamespace
    return ns(nerf=ns(sampling=ns(num_samples=NUM_SAMPLES),
                      propnet=ns(num_samples_per_prop=list(PROP_SAMPLES), near_plane=NEAR,
                                 far_plane=FAR, sampling_type="uniform_lindisp")),
              render=ns(render_chunk_size=16384))


def build_models(ns, case: str, seed: int = 0):
    """Build (field, propnets) from any namespace exposing the reference's class names
    (HashEncoder, RadianceField, build_density_field): the reference itself or the drop-in."""
    c = CASES[case]
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
        interpolate_xyz_encoding=True, enable_learnable_pe=True,
        enable_temporal_interpolation=False)
    field.register_normalized_training_timesteps(torch.linspace(0, 1, N_TIMESTEPS),
                                                 time_diff=1.0 / N_TIMESTEPS)
    props = []
    for e in ENC_PROP:
        p = ns.build_density_field(n_input_dims=3, n_levels=e["n_levels"],
                                   max_resolution=e["max_resolution"],
                                   log2_hashmap_size=e["log2_hashmap_size"],
                                   n_features_per_level=e["n_features_per_level"], unbounded=True)
        p.set_aabb(AABB)
        props.append(p)
    # tcnn-style U(-1e-4,1e-4) tables make every density ~exp(-1): rescale so the test scene
    # has structure (surfaces, non-uniform CDFs).  Deterministic given `seed`.
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for m in [field] + props:
            for k, v in m.named_parameters():
                if k.endswith("tcnn_encoding.params"):
                    v.copy_(torch.randn(v.shape, generator=g) * 0.5)
    return field, props


def make_batch(case: str, seed: int = 0, n_rays: int = N_RAYS, lidar: bool = False) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(1000 + seed)
    r = lambda *s: torch.rand(*s, generator=g)
    d = torch.randn(n_rays, 3, generator=g)
    d[:, 2] *= 0.3
    d = d / d.norm(dim=-1, keepdim=True)
    o = torch.stack([r(n_rays) * 60 - 10, r(n_rays) * 10 - 5, r(n_rays) * 2 + 1], -1)
    p = "lidar_" if lidar else ""
    out = {p + "origins": o, p + "viewdirs": d, p + "normed_timestamps": r(n_rays)}
    if lidar:
        out["lidar_ranges"] = r(n_rays, 1) * 70 + 1
    else:
        out["img_idx"] = torch.randint(0, N_TIMESTEPS * N_CAMS, (n_rays,), generator=g)
        out["pixel_coords"] = r(n_rays, 2)
        out["pixels"] = r(n_rays, 3)
        out["sky_masks"] = (r(n_rays) < 0.2).float()
        if CASES[case]["feature"]:
            out["features"] = r(n_rays, 64)
    return out
