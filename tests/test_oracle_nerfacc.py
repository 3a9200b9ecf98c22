"""Known-answer tests for the nerfacc restatement (the only published vectors available
offline: the ``importance_sampling`` docstring example of nerfacc@8340e19)."""
import torch

from oracle import nerfacc_ref as nf


def test_importance_sampling_docstring_example():
    # nerfacc/pdf.py docstring: ray 0 vals [0,1] cdfs [0,.5]; ray 1 vals [0,1,2] cdfs [0,.5,1]; n=2
    iv, sm = nf.importance_sampling(nf.RayIntervals(torch.tensor([[0.0, 1.0]])), torch.tensor([[0.0, 0.5]]), 2)
    assert torch.equal(iv.vals, torch.tensor([[0.0, 0.5, 1.0]]))
    assert torch.equal(sm.vals, torch.tensor([[0.25, 0.75]]))
    iv, sm = nf.importance_sampling(nf.RayIntervals(torch.tensor([[0.0, 1.0, 2.0]])),
                                    torch.tensor([[0.0, 0.5, 1.0]]), 2)
    assert torch.equal(iv.vals, torch.tensor([[0.0, 1.0, 2.0]]))
    assert torch.equal(sm.vals, torch.tensor([[0.5, 1.5]]))


def test_importance_sampling_edges_sorted_and_bounded():
    g = torch.Generator().manual_seed(0)
    R, m, n = 64, 33, 17
    vals = torch.sort(torch.rand(R, m, generator=g), -1).values
    w = torch.rand(R, m - 1, generator=g)
    cdfs = torch.cat([torch.zeros(R, 1), torch.cumsum(w / w.sum(-1, keepdim=True), -1)], -1)
    cdfs[:, -1] = 1.0
    for strat in (False, True):
        iv, _ = nf.importance_sampling(nf.RayIntervals(vals), cdfs, n, strat, jitter=torch.rand(R, 1, generator=g))
        e = iv.vals
        assert e.shape == (R, n + 1)
        assert (e[:, 1:] >= e[:, :-1]).all()
        assert (e >= vals[:, :1]).all() and (e <= vals[:, -1:]).all()


def test_volrend_identities():
    g = torch.Generator().manual_seed(1)
    t = torch.sort(torch.rand(8, 17, generator=g), -1).values
    sig = torch.rand(8, 16, generator=g) * 5
    w, tr, al = nf.render_weight_from_density(t[:, :-1], t[:, 1:], sig)
    # telescoping: sum of weights = 1 - T_end
    t_end = tr[:, -1] * (1 - al[:, -1])
    assert torch.allclose(w.sum(-1), 1 - t_end, atol=1e-6)
    assert torch.equal(tr[:, 0], torch.ones(8))
    assert torch.allclose(nf.accumulate_along_rays(w, None)[:, 0], w.sum(-1))
