"""The benchmark / parity configurations of BASELINE.json as attribute trees with the reference's
key names (configs/default_config.yaml, default_dynamic.yaml, default_flow.yaml), plus builders that
mirror builders.py:20-47,92-149 for the hot-path objects (field, proposal networks, estimator).
No OmegaConf needed: ``render_rays`` only does attribute access."""
from __future__ import annotations

import itertools
from types import SimpleNamespace as NS
from typing import List, Tuple

import torch

VARIANTS = {
    # name: (dynamic, shadow, flow, feature)            BASELINE.json configs[i]
    "static": (False, False, False, False),             # [1] default_config.yaml static field
    "dynamic": (True, True, False, False),              # [2] default_dynamic.yaml
    "flow": (True, True, True, False),                  # [3] default_flow.yaml
    "flow_feat": (True, True, True, True),              # [4] default_flow.yaml + feature head
}

AABB = [-20.0, -40.0, 0.0, 80.0, 40.0, 20.0]           # default_config.yaml:42


def make_cfg(variant: str = "static", num_timesteps: int = 200, num_cams: int = 3,
             num_samples: int = 64, prop_samples=(128, 64), small: bool = False) -> NS:
    """``small=True`` shrinks tables/samples for the CPU plumbing case (BASELINE configs[0])."""
    dynamic, shadow, flow, feature = VARIANTS[variant]
    log2 = 14 if small else 20
    log2_dyn = 13 if small else 18
    model = NS(
        xyz_encoder=NS(type="HashEncoder", n_input_dims=3, n_levels=10, n_features_per_level=4,
                       base_resolution=16, max_resolution=8192, log2_hashmap_size=log2),
        dynamic_xyz_encoder=NS(type="HashEncoder", n_input_dims=4, n_levels=10, n_features_per_level=4,
                               base_resolution=32, max_resolution=8192, log2_hashmap_size=log2_dyn),
        neck=NS(base_mlp_layer_width=64, geometry_feature_dim=64, semantic_feature_dim=64),
        head=NS(head_mlp_layer_width=64, enable_cam_embedding=False, enable_img_embedding=True,
                appearance_embedding_dim=16, enable_sky_head=True, enable_feature_head=feature,
                feature_embedding_dim=64, feature_mlp_layer_width=64, enable_learnable_pe=True,
                enable_dynamic_branch=dynamic, enable_shadow_head=shadow, interpolate_xyz_encoding=True,
                enable_temporal_interpolation=False, enable_flow_branch=flow),
        unbounded=True, num_cams=num_cams, num_train_timesteps=num_timesteps, resume_from=None,
    )
    nerf = NS(
        aabb=AABB, unbounded=True,
        propnet=NS(num_samples_per_prop=list(prop_samples), near_plane=0.1, far_plane=1000.0,
                   sampling_type="uniform_lindisp", enable_anti_aliasing_level_loss=True,
                   anti_aliasing_pulse_width=[0.03, 0.003],
                   xyz_encoder=NS(type="HashEncoder", n_input_dims=3, n_levels_per_prop=[8, 8],
                                  base_resolutions_per_prop=[16, 16], max_resolution_per_prop=[512, 2048],
                                  lgo2_hashmap_size_per_prop=[log2, log2], n_features_per_level=1)),
        sampling=NS(num_samples=num_samples), model=model)
    return NS(nerf=nerf, render=NS(render_chunk_size=16384),
              optim=NS(num_iters=25000, weight_decay=1e-5, lr=0.01, seed=0),
              data=NS(ray_batch_size=8192, num_timesteps=num_timesteps, num_cams=num_cams),
              variant=variant)


def build_hot_path(cfg: NS, device="cuda", table_std: float = 0.0, seed: int = 0, capturable: bool = False,
                   optimizer: str = "torch"):
    """(field, proposal_networks, estimator, optimizer) as builders.py builds them.  ``table_std > 0``
    replaces tcnn's degenerate U(-1e-4, 1e-4) table init by N(0, table_std) ("trained-like" state for
    bandwidth measurements).  ``optimizer``: "torch" = torch.optim.Adam exactly as builders.py:50-61
    constructs it, "fused" = emernerf_b200.optim.FusedAdam (same arithmetic, one launch per step, flat buffers)."""
    from .radiance_fields import build_density_field, build_radiance_field_from_cfg
    from .third_party.nerfacc_prop_net import PropNetEstimator

    torch.manual_seed(seed)
    m = cfg.nerf.model
    field = build_radiance_field_from_cfg(m, verbose=False)
    field.register_normalized_training_timesteps(
        torch.linspace(0, 1, m.num_train_timesteps), time_diff=1.0 / m.num_train_timesteps)
    field.set_aabb(cfg.nerf.aabb)
    pe = cfg.nerf.propnet.xyz_encoder
    props = []
    for i in range(len(cfg.nerf.propnet.num_samples_per_prop)):
        p = build_density_field(n_input_dims=pe.n_input_dims, n_levels=pe.n_levels_per_prop[i],
                                max_resolution=pe.max_resolution_per_prop[i],
                                log2_hashmap_size=pe.lgo2_hashmap_size_per_prop[i],
                                n_features_per_level=pe.n_features_per_level, unbounded=cfg.nerf.unbounded)
        p.set_aabb(cfg.nerf.aabb)
        props.append(p)
    if table_std > 0:
        g = torch.Generator().manual_seed(seed + 17)
        with torch.no_grad():
            for mod in [field] + props:
                for k, v in mod.named_parameters():
                    if k.endswith("tcnn_encoding.params"):
                        v.copy_(torch.randn(v.shape, generator=g) * table_std)
    field = field.to(device)
    props = [p.to(device) for p in props]
    adam = dict(lr=cfg.optim.lr, eps=1e-15, weight_decay=cfg.optim.weight_decay, betas=(0.9, 0.99))
    if optimizer == "fused":
        from .optim import FusedAdam

        def groups(mods):
            """Two flat buckets per optimizer: the hash tables (the big gradients, complete as soon as the grid
            scatter has run) first, everything else (MLPs, embedding, PE map) last -- the order DataParallel reduces in."""
            named = [(k, v) for m in mods for k, v in m.named_parameters()]
            tables = [v for k, v in named if k.endswith("tcnn_encoding.params")]
            rest = [v for k, v in named if not k.endswith("tcnn_encoding.params")]
            return [{"params": tables}, {"params": rest}]

        prop_opt = FusedAdam(groups(props), flatten_params=True, **adam)
        est = PropNetEstimator(prop_opt, None,
                               enable_anti_aliasing_loss=cfg.nerf.propnet.enable_anti_aliasing_level_loss,
                               anti_aliasing_pulse_width=cfg.nerf.propnet.anti_aliasing_pulse_width).to(device)
        return field, props, est, FusedAdam(groups([field]), flatten_params=True, **adam)
    if capturable:                      # CUDA-graph capture of the optimizer step
        adam["capturable"] = True
    if str(device).startswith("cuda"):  # one fused kernel per step instead of the foreach chain (same maths)
        adam["fused"] = True
    prop_opt = torch.optim.Adam(itertools.chain(*[p.parameters() for p in props]), **adam)
    est = PropNetEstimator(prop_opt, None,
                           enable_anti_aliasing_loss=cfg.nerf.propnet.enable_anti_aliasing_level_loss,
                           anti_aliasing_pulse_width=cfg.nerf.propnet.anti_aliasing_pulse_width).to(device)
    opt = torch.optim.Adam(field.parameters(), **adam)
    return field, props, est, opt
