"""Host-side bookkeeping of the MLP head chain (emernerf_b200/_ops.py:_MLPChain) on CPU.

The three per-layer kernel launchers are replaced by torch stand-ins (this file only -- the product has
no CPU path), so what is checked here is the chain's own logic: which buffers the layers read and write,
the shared skip-concatenation buffer, the stacked skip gradient [dZ0 | dZ1] [W0 ; W1[:, h:]], and the
order of the returned gradients, against autograd on a plain torch MLP with the reference's
concatenation order [hidden | input] (radiance_fields/mlp.py:38-46).
"""
import pytest
import torch

from emernerf_b200 import _ops


def _act(v, act):
    if act == _ops.ACT_RELU:
        return torch.relu(v)
    if act == _ops.ACT_SIGMOID:
        return torch.sigmoid(v)
    return v


@pytest.fixture
def cpu_layers(monkeypatch):
    calls = []

    def fwd(x2, ldx, w, b, y, ldy, n, act):
        assert x2.stride(0) == ldx or n == 1
        assert y.stride(0) == ldy or n == 1
        calls.append(("fwd", tuple(w.shape)))
        v = x2[:, :w.shape[1]] @ w.t()
        if b is not None:
            v = v + b
        y.copy_(_act(v, act))

    def bwd_data(dz, lddz, w, dx, lddx, n, relu_src, ld_relu, relu_cols):
        n_out, k = w.shape
        assert dz.stride(0) == lddz or n == 1
        assert dx.stride(0) == lddx or n == 1
        calls.append(("bwd_data", (n_out, k)))
        g = dz[:, :n_out] @ w
        if relu_src is not None:
            g[:, :relu_cols] = g[:, :relu_cols] * (relu_src[:, :relu_cols] > 0)
        dx[:, :k] = g

    def bwd_weight(x2, ldx, dz, lddz, w, has_bias, n, w_sink=None, b_sink=None):
        n_out, k = w.shape
        calls.append(("bwd_weight", (n_out, k)))
        return dz[:, :n_out].t() @ x2[:, :k], (dz[:, :n_out].sum(0) if has_bias else None)

    monkeypatch.setattr(_ops, "_need_cuda", lambda *ts: None)
    monkeypatch.setattr(_ops, "_layer_fwd", fwd)
    monkeypatch.setattr(_ops, "_layer_bwd_data", bwd_data)
    monkeypatch.setattr(_ops, "_layer_bwd_weight", bwd_weight)
    return calls


def _reference(x, ws, bs, out_act, skip):
    h = x
    for i, (w, b) in enumerate(zip(ws, bs)):
        if i == skip and i > 0:
            h = torch.cat([h, x], -1)
        h = h @ w.t() + (b if b is not None else 0.0)
        h = torch.relu(h) if i < len(ws) - 1 else _act(h, out_act)
    return h


def _params(dims, k0, skip, bias, seed):
    g = torch.Generator().manual_seed(seed)
    ws, bs = [], []
    k = k0
    for i, d in enumerate(dims):
        if i == skip and i > 0:
            k += k0
        ws.append((torch.randn(d, k, generator=g, dtype=torch.float64) * 0.4).float().requires_grad_())
        bs.append((torch.randn(d, generator=g, dtype=torch.float64) * 0.1).float().requires_grad_() if bias else None)
        k = d
    return ws, bs


CASES = [
    # dims, k0, skip, bias, out_act
    ((8, 8, 3), 13, 1, True, _ops.ACT_SIGMOID),      # the colour head's shape: stacked skip gradient
    ((8, 12, 5, 3), 6, 1, True, _ops.ACT_NONE),      # deeper chain, stacked
    ((8, 3), 13, 1, True, _ops.ACT_NONE),            # skip layer is the last one: two products + add
    ((6, 8, 3), 13, 1, False, _ops.ACT_RELU),        # hidden width not a multiple of 4: two products + add
    ((8, 8, 3), 13, 2, True, _ops.ACT_NONE),         # skip further down
    ((8, 8, 3), 13, -1, True, _ops.ACT_SIGMOID),     # no skip
    ((5,), 7, -1, True, _ops.ACT_NONE),              # single layer
]


@pytest.mark.parametrize("impl", ["stack", "add"])
@pytest.mark.parametrize("shared", [False, True])
@pytest.mark.parametrize("dims,k0,skip,bias,out_act", CASES)
def test_chain_matches_autograd(cpu_layers, monkeypatch, impl, shared, dims, k0, skip, bias, out_act):
    monkeypatch.setattr(_ops, "SKIP_BWD_IMPL", impl)
    ws, bs = _params(dims, k0, skip, bias, seed=3)
    g = torch.Generator().manual_seed(11)
    r, s_ = 5, 4
    x_val = torch.randn(r, s_, k0, generator=g)
    catbuf = None
    if shared and skip > 0:
        h = dims[skip - 1]
        catbuf = torch.full((r * s_, _ops._pad4(h + k0)), float("nan"))
        catbuf[:, h + k0:] = 0.0
        catbuf[:, h:h + k0] = x_val.reshape(-1, k0)
        # the chain must be handed the VIEW of catbuf itself (as field_tail does), with a leaf to collect dX
        leaf = torch.zeros(r, s_, k0, requires_grad=True)
        x = _Alias.apply(leaf, catbuf[:, h:h + k0].view(r, s_, k0))
    else:
        leaf = x_val.clone().requires_grad_()
        x = leaf
    y = _ops.mlp_chain(x, ws, bs if bias else None, out_act, skip, catbuf=catbuf)
    up = torch.randn(y.shape, generator=g)
    (y * up).sum().backward()

    ws_r = [w.detach().clone().requires_grad_() for w in ws]
    bs_r = [None if b is None else b.detach().clone().requires_grad_() for b in bs]
    x_r = x_val.clone().requires_grad_()
    y_r = _reference(x_r, ws_r, bs_r, out_act, skip)
    (y_r * up).sum().backward()

    torch.testing.assert_close(y, y_r, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(leaf.grad, x_r.grad, rtol=1e-5, atol=1e-6)
    for w, w_r in zip(ws, ws_r):
        torch.testing.assert_close(w.grad, w_r.grad, rtol=1e-5, atol=1e-6)
    for b, b_r in zip(bs, bs_r):
        if b is not None:
            torch.testing.assert_close(b.grad, b_r.grad, rtol=1e-5, atol=1e-6)

    if shared and skip > 0 and skip < len(dims):
        h = dims[skip - 1]
        # the layer in front of the skip wrote straight into the caller's buffer; nothing was copied
        assert not torch.isnan(catbuf[:, :h + k0]).any()
    stacked = (impl == "stack" and skip == 1 and len(dims) >= 3 and dims[0] % 4 == 0 and dims[1] % 4 == 0)
    shapes = [c[1] for c in cpu_layers if c[0] == "bwd_data"]
    if stacked:
        assert (dims[0] + dims[1], k0) in shapes and (dims[1], dims[0]) in shapes
        assert (dims[1], dims[0] + k0) not in shapes
    elif skip > 0 and skip < len(dims):
        assert (dims[skip], dims[skip - 1] + k0) in shapes


class _Alias(torch.autograd.Function):
    """Hands ``view`` (a tensor aliasing caller-owned memory) to the graph with ``leaf`` as its gradient sink."""

    @staticmethod
    def forward(ctx, leaf, view):
        return view.view_as(view)

    @staticmethod
    def backward(ctx, g):
        return g, None
