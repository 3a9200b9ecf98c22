"""Hash-grid restatement: level tables and structural properties."""
import torch

from oracle import hotpath, tcnn_ref

TABLE = {  # name: (D, HashEncoder args, entries, resolutions of first/last level, n dense levels)
    "static": (3, (10, 16, 8192, 20, 4), 7_639_040, (16, 8192), 3),
    "dynamic": (4, (10, 32, 8192, 18, 4), 2_621_440, (32, 8192), 0),
    "flow": (4, (10, 16, 4096, 18, 4), 2_424_832, (16, 4096), 1),
    "prop0": (3, (8, 16, 512, 20, 1), 4_661_184, (16, 512), 4),
    "prop1": (3, (8, 16, 2048, 20, 1), 5_541_888, (16, 2048), 3),
}


def test_level_tables_match_survey():
    for name, (D, args, entries, (r0, r1), n_dense) in TABLE.items():
        g = tcnn_ref.grid_geometry(D, hotpath.hash_encoder_config(*args))
        assert g.offsets[-1] == entries, name
        assert (g.resolutions[0], g.resolutions[-1]) == (r0, r1), name
        assert sum(not h for h in g.hashed) == n_dense, name
        assert all(o % 8 == 0 for o in g.offsets), name


def test_weights_partition_of_unity_and_linear_reproduction():
    cfg = hotpath.hash_encoder_config(3, 4, 16, 12, 2)   # all levels dense (res<=16^3<=4096)
    g = tcnn_ref.grid_geometry(3, cfg)
    assert not any(g.hashed)
    x = torch.rand(257, 3) * 0.7 + 0.05      # keep cell+1 < res at level 0 (no dense wrap-around)
    for lvl in range(g.n_levels):
        idx, w, frac, cell = tcnn_ref.corner_indices_and_weights(x, g, lvl)
        assert torch.allclose(w.sum(-1), torch.ones(257), atol=1e-6)
        assert (idx >= g.offsets[lvl]).all() and (idx < g.offsets[lvl + 1]).all()
    # a table that stores an affine function of the vertex coordinate is reproduced exactly
    lvl = 0
    res, scale = g.resolutions[0], g.scales[0]
    params = torch.zeros(g.n_params)
    table = params.view(-1, 2)
    ii = torch.arange(res)
    vx, vy, vz = torch.meshgrid(ii, ii, ii, indexing="ij")
    lin = (vx + 2 * vy + 3 * vz).float()
    flat_idx = (vx + vy * res + vz * res * res).reshape(-1)
    table[g.offsets[0] + flat_idx, 0] = lin.reshape(-1)
    y = tcnn_ref.grid_forward(x, params, g)[:, 0]
    pos = x * scale + 0.5
    want = pos[:, 0] + 2 * pos[:, 1] + 3 * pos[:, 2]
    assert torch.allclose(y, want, atol=2e-4)


def test_input_gradient_is_scale_times_finite_difference():
    cfg = hotpath.hash_encoder_config(2, 8, 16, 10, 4)
    g = tcnn_ref.grid_geometry(4, cfg)
    params = torch.randn(g.n_params, dtype=torch.float32)
    x = (torch.rand(33, 4) * 0.8 + 0.1).requires_grad_(True)
    y = tcnn_ref.grid_forward(x, params, g)
    (gx,) = torch.autograd.grad(y.sum(), x)
    eps = 1e-4
    for d in range(4):
        xp = x.detach().clone(); xp[:, d] += eps
        xm = x.detach().clone(); xm[:, d] -= eps
        fd = (tcnn_ref.grid_forward(xp, params, g).sum(-1) - tcnn_ref.grid_forward(xm, params, g).sum(-1)) / (2 * eps)
        ok = (fd - gx[:, d]).abs() < 5e-2 * gx[:, d].abs().clamp_min(1.0)
        assert ok.float().mean() > 0.9     # cells crossed by the +-eps stencil are the exceptions
