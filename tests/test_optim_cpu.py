"""FusedAdam's host logic on CPU (C ABI answered by tests/cabi_emulator.py): flat state, gradient sinks fed by the
library's backward functions, skipped untouched parameters, torch.optim.Adam's arithmetic and checkpoint layout."""
import copy
import types

import pytest
import torch

import cabi_emulator
import cases
from helpers import Golden, rel_err
from oracle import adapters

ADAM = dict(lr=0.01, eps=1e-15, weight_decay=1e-5, betas=(0.9, 0.99))        # builders.py:50-61


def _models(case="static"):
    from emernerf_b200.radiance_fields import RadianceField, build_density_field
    from emernerf_b200.radiance_fields.encodings import HashEncoder

    ns = types.SimpleNamespace(HashEncoder=HashEncoder, RadianceField=RadianceField, build_density_field=build_density_field)
    field, props = cases.build_models(ns, case)
    g = Golden(case)
    field.load_state_dict(g.tensors("sd/field"))
    [p.load_state_dict(g.tensors(f"sd/prop{i}")) for i, p in enumerate(props)]
    return g, field, props


def _loss(g, field, props, est):
    from emernerf_b200.radiance_fields.render_utils import render_rays

    field.train(); est.train()
    [p.train() for p in props]
    est._jitter_override, field._noise_override = g.jitters("train"), g.noise("train")
    est.prop_cache.clear()
    out = render_rays(field, est, props, g.tensors("in/pixel"), cases.render_cfg(), proposal_requires_grad=True)
    return adapters.parity_loss(out), est.compute_loss(out["extras"]["trans"], 1024.0)


@pytest.mark.parametrize("flatten", [False, True])
def test_fused_adam_follows_torch_adam_through_training_steps(flatten, monkeypatch):
    from emernerf_b200 import _ops
    from emernerf_b200.optim import FusedAdam
    from emernerf_b200.third_party.nerfacc_prop_net import PropNetEstimator

    cabi_emulator.install(monkeypatch)
    _ops.clear_grad_sinks()
    g, field_a, props_a = _models()
    field_b, props_b = copy.deepcopy(field_a), copy.deepcopy(props_a)
    prop_params = lambda ps: [q for m in ps for q in m.parameters()]
    opt_a, popt_a = torch.optim.Adam(field_a.parameters(), **ADAM), torch.optim.Adam(prop_params(props_a), **ADAM)
    opt_b = FusedAdam(field_b.parameters(), flatten_params=flatten, **ADAM)
    popt_b = FusedAdam(prop_params(props_b), flatten_params=flatten, **ADAM)
    est_a, est_b = PropNetEstimator(popt_a, None), PropNetEstimator(popt_b, None)
    before = cabi_emulator.CALLS.count("emer_adam_step")
    for step in range(3):
        for field, props, est, opt, popt in ((field_a, props_a, est_a, opt_a, popt_a), (field_b, props_b, est_b, opt_b, popt_b)):
            loss, ploss = _loss(g, field, props, est)
            popt.zero_grad()
            ploss.backward()
            popt.step()
            opt.zero_grad()
            (loss * 1024.0).backward()                  # the reference's never-unscaled GradScaler (Q17)
            opt.step()
    assert cabi_emulator.CALLS.count("emer_adam_step") - before == 6          # one launch per optimizer step
    # (eps = 1e-15 makes the update m / sqrt(v): scale-free, so where a gradient is rounding noise -- the sky head's is
    # proportional to 1 - opacity, which is 0 or one ulp in this dense scene -- two runs that differ in the last bit
    # take different +-lr steps; those parameters are left out, the others are compared at 5e-4 and the arithmetic
    # itself at 1e-6 in the next test)
    for (k, a), (_, b) in zip(field_a.named_parameters(), field_b.named_parameters()):
        if "sky_head" in k:
            continue
        assert rel_err(b, a) < 5e-4, (k, rel_err(b, a))
    for pa, pb in zip(props_a, props_b):
        for (k, a), (_, b) in zip(pa.named_parameters(), pb.named_parameters()):
            assert rel_err(b, a) < 5e-4, (k, rel_err(b, a))
    # proposal network 0 is never evaluated (Q21): no gradient, so -- as with torch -- no update, weight decay included
    g0 = Golden("static").tensors("sd/prop0")
    for k, v in props_b[0].state_dict().items():
        assert torch.equal(v, g0[k]), k
    # gradients were consumed AND zeroed by the step
    assert all(float(f.abs().max()) == 0.0 for f in opt_b.flat_grads() + popt_b.flat_grads())
    # the big consumers accumulated straight into the sinks: autograd never saw a table gradient
    table = field_b.xyz_encoder.tcnn_encoding.params
    assert table.grad.data_ptr() == opt_b.flat_grads()[0].data_ptr() + 4 * opt_b._groups[0].offsets[
        [id(p) for p in opt_b._groups[0].params].index(id(table))]

    # checkpoints: torch.optim.Adam's layout both ways
    sd = opt_b.state_dict()
    opt_c = torch.optim.Adam(field_b.parameters(), **ADAM)
    opt_c.load_state_dict(sd)
    some = next(iter(opt_c.state.values()))
    assert float(some["step"]) == 3.0 and set(some) >= {"step", "exp_avg", "exp_avg_sq"}
    _ops.clear_grad_sinks()
    g2, field_d, _ = _models()
    opt_d = FusedAdam(field_d.parameters(), **ADAM)
    opt_d.load_state_dict(opt_a.state_dict())
    for pa, pd in zip(field_a.parameters(), field_d.parameters()):
        assert torch.equal(opt_a.state[pa]["exp_avg"], opt_d.state[pd]["exp_avg"])
    assert float(opt_d._groups[0].hyper[0]) == 3.0
    _ops.clear_grad_sinks()


def test_fused_adam_arithmetic_matches_torch(monkeypatch):
    """Same gradients in, same parameters out (<= 1e-6 relative after 4 steps), lr change between steps included."""
    from emernerf_b200 import _ops
    from emernerf_b200.optim import FusedAdam

    cabi_emulator.install(monkeypatch)
    _ops.clear_grad_sinks()
    torch.manual_seed(1)
    shapes = ((257, 9), (64,), (3, 64), (4099,))
    pa = [torch.nn.Parameter(torch.randn(s)) for s in shapes]
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    a, b = torch.optim.Adam(pa, **ADAM), FusedAdam(pb, **ADAM)
    for step in range(4):
        if step == 2:
            for o in (a, b):
                o.param_groups[0]["lr"] = 0.003
        for x, y in zip(pa, pb):
            gr = torch.randn_like(x) * (10.0 ** (step - 2))
            x.grad = gr.clone()
            y.grad.add_(gr)                              # what AccumulateGrad does with a defined .grad
        b.mark_all_touched()
        a.step(); b.step()
    for x, y in zip(pa, pb):
        assert rel_err(y, x) < 1e-6, rel_err(y, x)
    _ops.clear_grad_sinks()


def test_fused_adam_shard_updates_only_its_slice(monkeypatch):
    """``shard = (rank, world)``: the step touches this rank's slice of the flat space only (what the sharded data-parallel
    step runs between reduce-scatter and all-gather) and still clears the whole gradient buffer."""
    from emernerf_b200 import _ops
    from emernerf_b200.optim import ALIGN, FusedAdam

    cabi_emulator.install(monkeypatch)
    _ops.clear_grad_sinks()
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(s)) for s in ((300, 7), (64,), (5, 40), (1000,))]
    full = FusedAdam([torch.nn.Parameter(p.detach().clone()) for p in ps], flatten_params=True, **ADAM)
    parts = [FusedAdam([torch.nn.Parameter(p.detach().clone()) for p in ps], flatten_params=True, **ADAM) for _ in range(3)]
    grads = [torch.randn_like(p) for p in ps]
    for opt in [full] + parts:
        for p, gr in zip(opt._groups[0].params, grads):
            p.grad.copy_(gr)
        opt.mark_all_touched()
    full.step()
    want = full.flat_params()[0]
    got = torch.zeros_like(want)
    for r, opt in enumerate(parts):
        opt.shard = (r, 3)
        before = opt.flat_params()[0].clone()
        opt.step()
        tot = opt._groups[0].total
        per = tot // 3 if tot % (ALIGN * 3) == 0 else (tot // ALIGN + 2) // 3 * ALIGN
        lo, hi = min(r * per, opt._groups[0].total), min((r + 1) * per, opt._groups[0].total)
        after = opt.flat_params()[0]
        assert torch.equal(after[:lo], before[:lo]) and torch.equal(after[hi:], before[hi:])
        got[lo:hi] = after[lo:hi]
        assert float(opt.flat_grads()[0].abs().max()) == 0.0
    assert torch.equal(got, want)
    _ops.clear_grad_sinks()
