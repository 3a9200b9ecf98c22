"""Option variants of the field / estimator, REFERENCE vs DROP-IN, live (build container only).

Run as a script by tests/test_live_reference_variants.py in a subprocess (the reference's module names
``radiance_fields`` / ``third_party`` go into ``sys.modules``).  For every variant the reference's own classes
(with the oracle stand-ins for tiny-cuda-nn / nerfacc, oracle/ref_shims.py) and the drop-in's classes (with the
C ABI answered by tests/cabi_emulator.py) are built with the same constructor arguments, the reference's
state-dict is loaded into the drop-in, both render the same rays, and the maximum relative error of every
output is printed as one JSON object.  What this covers beyond the four golden cases: the constructor options
and data-dict shapes of SURVEY.md section 8b that the shipped configs do not exercise.
"""
from __future__ import annotations

import json
import os
import sys
import types
import warnings

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE, os.path.join(HERE, "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)
warnings.filterwarnings("ignore")

from oracle import ref_shims  # noqa: E402

ref_shims.install()

import cabi_emulator  # noqa: E402
import cases  # noqa: E402
from radiance_fields import RadianceField as RefField, build_density_field as ref_build_density  # noqa: E402
from radiance_fields.encodings import HashEncoder as RefEncoder  # noqa: E402
from radiance_fields.render_utils import render_rays as ref_render_rays  # noqa: E402
from third_party.nerfacc_prop_net import PropNetEstimator as RefEstimator  # noqa: E402

from emernerf_b200.radiance_fields import RadianceField, build_density_field  # noqa: E402
from emernerf_b200.radiance_fields.encodings import HashEncoder  # noqa: E402
from emernerf_b200.radiance_fields.render_utils import render_rays  # noqa: E402
from emernerf_b200.third_party.nerfacc_prop_net import PropNetEstimator  # noqa: E402


class _Patch:
    """monkeypatch-like object for cabi_emulator.install outside pytest."""

    def setattr(self, obj, name, value):
        setattr(obj, name, value)


cabi_emulator.install(_Patch())

REF = types.SimpleNamespace(HashEncoder=RefEncoder, RadianceField=RefField, build_density_field=ref_build_density)
OURS = types.SimpleNamespace(HashEncoder=HashEncoder, RadianceField=RadianceField,
                             build_density_field=build_density_field)

BASE = dict(geometry_feature_dim=64, base_mlp_layer_width=64, head_mlp_layer_width=64, enable_cam_embedding=False,
            enable_img_embedding=True, num_cams=cases.N_CAMS, appearance_embedding_dim=16, semantic_feature_dim=64,
            feature_mlp_layer_width=64, feature_embedding_dim=64, enable_sky_head=True, enable_shadow_head=False,
            enable_feature_head=False, num_train_timesteps=cases.N_TIMESTEPS, interpolate_xyz_encoding=True,
            enable_learnable_pe=True, enable_temporal_interpolation=False, unbounded=True)

# name -> (field kwargs, dynamic grid?, flow grid?, batch edits, render cfg edits, estimator kwargs, mode)
VARIANTS = {
    "cam_embedding": (dict(enable_cam_embedding=True, enable_img_embedding=False), False, False, "cam_idx", {}, {}, "eval"),
    "no_embedding": (dict(enable_img_embedding=False), False, False, None, {}, {}, "eval"),
    "mean_embedding": ({}, False, False, "drop_idx", {}, {}, "eval"),          # novel view: no img_idx in the batch
    "bounded_aabb": (dict(unbounded=False), False, False, None, {}, {}, "eval"),
    "no_sky_head": (dict(enable_sky_head=False), False, False, None, {}, {}, "eval"),
    "narrow_widths": (dict(geometry_feature_dim=32, base_mlp_layer_width=32, head_mlp_layer_width=32,
                           appearance_embedding_dim=8), False, False, None, {}, {}, "eval"),
    "odd_geometry_width": (dict(geometry_feature_dim=15), False, False, None, {}, {}, "eval"),   # the class default
    "dynamic_no_shadow": ({}, True, False, None, {}, {}, "eval"),
    "feature_head_no_pe": (dict(enable_feature_head=True, enable_learnable_pe=False), True, True, "features", {}, {},
                           "eval"),
    "wide_embedding": (dict(appearance_embedding_dim=48), False, False, None, {}, {}, "train"),   # > 32: generic tail
    "wide_heads": (dict(head_mlp_layer_width=256, base_mlp_layer_width=256, geometry_feature_dim=128), True, False,
                   None, {}, {}, "train"),                 # layers the tensor-core kernels cannot hold -> CUDA-core path
    "wide_feature_head": (dict(enable_feature_head=True, feature_mlp_layer_width=256, feature_embedding_dim=384,
                               semantic_feature_dim=32), True, True, "features384", {}, {}, "train"),
    "dynamic_model_no_time": ({}, True, False, "drop_time", {}, {}, "eval"),      # no timestamps: static branch only
    "one_proposal": ({}, False, False, None, dict(num_samples_per_prop=[24], n_props=1), {}, "train"),
    "three_proposals": ({}, False, False, None, dict(num_samples_per_prop=[40, 24, 16], n_props=3), {}, "eval"),
    "many_samples": ({}, False, False, None, dict(num_samples_per_prop=[300, 280], num_samples=270), {}, "eval"),
    "sampling_lindisp": ({}, False, False, None, dict(sampling_type="lindisp"), {}, "eval"),
    "sampling_uniform": ({}, False, False, None, dict(sampling_type="uniform", far_plane=120.0), {}, "eval"),
    "train_stratified": ({}, True, False, None, {}, {}, "train"),
    "train_plain_pdf_loss": ({}, False, False, None, {}, dict(enable_anti_aliasing_loss=False), "train"),
}


def build(ns, kwargs, dynamic, flow, seed=0):
    torch.manual_seed(seed)
    enc = ns.HashEncoder(verbose=False, **cases.ENC_STATIC)
    dyn = ns.HashEncoder(verbose=False, **cases.ENC_DYN) if dynamic else None
    flw = ns.HashEncoder(verbose=False, **cases.ENC_FLOW) if flow else None
    kw = dict(BASE)
    kw.update(kwargs)
    field = ns.RadianceField(xyz_encoder=enc, dynamic_xyz_encoder=dyn, flow_xyz_encoder=flw, aabb=cases.AABB, **kw)
    field.register_normalized_training_timesteps(torch.linspace(0, 1, cases.N_TIMESTEPS),
                                                 time_diff=1.0 / cases.N_TIMESTEPS)
    props = []
    for e in cases.ENC_PROP:
        p = ns.build_density_field(n_input_dims=3, n_levels=e["n_levels"], max_resolution=e["max_resolution"],
                                   log2_hashmap_size=e["log2_hashmap_size"],
                                   n_features_per_level=e["n_features_per_level"], unbounded=kw["unbounded"])
        p.set_aabb(cases.AABB)
        props.append(p)
    return field, props


def randomise(field, props, seed=1):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in [field] + props:
            for k, v in m.named_parameters():
                if k.endswith("tcnn_encoding.params"):
                    v.copy_(torch.randn(v.shape, generator=g) * 0.5)


def rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def compare(got, want, errs, prefix=""):
    assert set(got) == set(want), (prefix, sorted(set(got) ^ set(want)))
    for k in want:
        if isinstance(want[k], dict):
            compare(got[k], want[k], errs, prefix + k + "/")
        else:
            assert got[k].shape == want[k].shape, (prefix + k, got[k].shape, want[k].shape)
            errs[prefix + k] = rel(got[k], want[k])


def run_variant(name):
    kwargs, dynamic, flow, edit, cfg_edit, est_kw, mode = VARIANTS[name]
    rf, rp = build(REF, kwargs, dynamic, flow)
    randomise(rf, rp)
    of, op = build(OURS, kwargs, dynamic, flow)
    of.load_state_dict(rf.state_dict())
    for a, b in zip(op, rp):
        a.load_state_dict(b.state_dict())
    case = "flow_feat" if edit in ("features", "features384") else "static"
    batch = cases.make_batch(case)
    if edit == "features384":
        batch["features"] = torch.rand(batch["origins"].shape[0], 384, generator=torch.Generator().manual_seed(5))
    if edit == "cam_idx":
        batch["cam_idx"] = batch.pop("img_idx") % cases.N_CAMS
    elif edit == "drop_idx":
        batch.pop("img_idx")
    elif edit == "drop_time":
        batch.pop("normed_timestamps")
    cfg = cases.render_cfg()
    cfg_edit = dict(cfg_edit)
    n_props = cfg_edit.pop("n_props", None)
    if "num_samples" in cfg_edit:
        cfg.nerf.sampling.num_samples = cfg_edit.pop("num_samples")
    for k, v in cfg_edit.items():
        setattr(cfg.nerf.propnet, k, v)
    if n_props is not None:
        # the same networks on both sides: drop one, or append a third built like the second
        if n_props < len(rp):
            rp, op = rp[:n_props], op[:n_props]
        while len(rp) < n_props:
            e = cases.ENC_PROP[-1]
            extra = []
            for ns in (REF, OURS):
                torch.manual_seed(9)
                p = ns.build_density_field(n_input_dims=3, n_levels=e["n_levels"], max_resolution=e["max_resolution"],
                                           log2_hashmap_size=e["log2_hashmap_size"],
                                           n_features_per_level=e["n_features_per_level"], unbounded=True)
                p.set_aabb(cases.AABB)
                extra.append(p)
            randomise(extra[0], [], seed=3)
            extra[1].load_state_dict(extra[0].state_dict())
            rp, op = rp + [extra[0]], op + [extra[1]]
    train = mode == "train"
    r_est = RefEstimator(torch.optim.Adam([q for p in rp for q in p.parameters()], lr=0.01), None, **est_kw)
    o_est = PropNetEstimator(torch.optim.Adam([q for p in op for q in p.parameters()], lr=0.01), None, **est_kw)
    for m in (rf, of, r_est, o_est, *rp, *op):
        m.train(train)
    errs = {}
    failures = []
    with torch.set_grad_enabled(train):
        for fn, args in ((ref_render_rays, (rf, r_est, rp)), (render_rays, (of, o_est, op))):
            torch.manual_seed(77)
            try:
                failures.append(None)
                res = fn(*args, dict(batch), cfg, proposal_requires_grad=train, return_decomposition=not train)
            except (AssertionError, ValueError, KeyError) as e:          # same refusal on both sides is parity too
                failures[-1] = (type(e).__name__, str(e))
                res = None
            if fn is ref_render_rays:
                want = res
            else:
                got = res
    if failures[0] is not None or failures[1] is not None:
        assert failures[0] == failures[1], failures
        return {"raises:" + failures[0][0]: 0.0}
    compare(got, want, errs)
    if train:
        wl = r_est.compute_loss(want["extras"]["trans"], 1024.0)
        gl = o_est.compute_loss(got["extras"]["trans"], 1024.0)
        errs["prop_loss"] = abs(gl.item() - wl.item()) / max(1.0, abs(wl.item()))
        loss_w = (want["rgb"] - batch["pixels"]).square().mean() + want["depth"].mean() * 1e-2
        loss_g = (got["rgb"] - batch["pixels"]).square().mean() + got["depth"].mean() * 1e-2
        (loss_w + wl).backward()
        (loss_g + gl).backward()
        ref_grads = dict(rf.named_parameters())
        for k, v in of.named_parameters():
            w = ref_grads[k].grad
            assert (v.grad is None) == (w is None), k
            if w is not None:
                errs["grad/" + k] = rel(v.grad, w)
        for i, (a, b) in enumerate(zip(op, rp)):
            rg = dict(b.named_parameters())
            for k, v in a.named_parameters():
                w = rg[k].grad
                assert (v.grad is None) == (w is None), (i, k)
                if w is not None:
                    errs[f"grad/prop{i}/" + k] = rel(v.grad, w)
    return errs


if __name__ == "__main__":
    assert ref_shims.reference_available(), "needs /root/reference"
    out = {}
    for name in (sys.argv[1:] or list(VARIANTS)):
        del cabi_emulator.CALLS[:]
        out[name] = {"errors": run_variant(name), "calls": sorted(set(cabi_emulator.CALLS))}
    print("JSON:" + json.dumps(out))
