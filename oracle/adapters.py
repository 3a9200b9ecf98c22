"""Glue between nn.Module-style fields (the reference's or the drop-in's -- same attribute
names) and the functional oracle.  TEST INFRASTRUCTURE -- see ``oracle/__init__.py``."""
from __future__ import annotations

from typing import Dict

import torch

from .hotpath import FieldSpec


def _enc_cfg(enc):
    return None if enc is None else dict(enc.encoding_config)


def spec_from_module(m) -> FieldSpec:
    """Read a RadianceField / DensityField (radiance_field.py:21-217,788-812) into a FieldSpec."""
    if not hasattr(m, "rgb_head"):          # DensityField
        return FieldSpec(xyz=_enc_cfg(m.xyz_encoder), unbounded=bool(m.unbounded), density_only=True)
    td = getattr(m, "time_diff", 0.0)
    return FieldSpec(
        xyz=_enc_cfg(m.xyz_encoder),
        dynamic=_enc_cfg(getattr(m, "dynamic_xyz_encoder", None)),
        flow=_enc_cfg(getattr(m, "flow_xyz_encoder", None)),
        unbounded=bool(m.unbounded),
        geometry_feature_dim=int(m.geometry_feature_dim),
        semantic_feature_dim=int(m.semantic_feature_dim),
        enable_cam_embedding=bool(m.enable_cam_embedding),
        enable_img_embedding=bool(m.enable_img_embedding),
        appearance_embedding_dim=int(m.appearance_embedding_dim),
        enable_sky_head=bool(m.enable_sky_head),
        enable_shadow_head=bool(m.enable_shadow_head),
        enable_feature_head=bool(m.enable_feature_head),
        enable_learnable_pe=bool(getattr(m, "enable_learnable_pe", True)),
        time_diff=float(td),
    )


def cpu_state_dict(m, requires_grad: bool = False) -> Dict[str, torch.Tensor]:
    sd = {k: v.detach().to("cpu", copy=True) for k, v in m.state_dict().items()}
    if requires_grad:
        for k, v in sd.items():
            if v.dtype.is_floating_point and k not in ("aabb", "training_timesteps",
                    "feats_reduction_mat", "feat_color_min", "feat_color_max",
                    "direction_encoding.scales"):
                v.requires_grad_(True)
    return sd


def parity_loss(out: dict) -> torch.Tensor:
    """A fixed scalar that touches every differentiable output of render_rays (used to compare
    parameter gradients between implementations; not one of the reference's training losses)."""
    ex = out["extras"]
    loss = out["depth"].mean() * 0.01 + out["opacity"].mean() + ex["weights"].square().mean()
    for k, c in (("rgb", 1.0), ("dino_feat", 0.5), ("shadow_ratio", 0.3)):
        if k in out:
            loss = loss + c * out[k].square().mean()
    if "dynamic_density" in ex:
        loss = loss + 0.01 * ex["dynamic_density"].mean()
    if "forward_pred_backward_flow" in ex:      # train_emernerf.py:700-716 cycle loss
        loss = loss + 0.5 * ((ex["forward_flow"].detach() + ex["forward_pred_backward_flow"]) ** 2
                             + (ex["backward_flow"].detach() + ex["backward_pred_forward_flow"]) ** 2).mean()
    return loss
