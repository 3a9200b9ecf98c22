"""``MLP`` with input-skip concatenation (radiance_fields/mlp.py:7-46 of the reference): same
constructor, same ``layers.{i}.{weight,bias}`` parameters; every layer runs on the library's
dense-layer kernels (bias + ReLU fused into the layer epilogue)."""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.nn as nn
from torch import Tensor

from .. import _ops


class MLP(nn.Module):
    def __init__(self, in_dims: int, out_dims: int, num_layers: int = 3, hidden_dims: Optional[int] = 256,
                 skip_connections: Optional[Sequence[int]] = (0,)) -> None:
        super().__init__()
        self.in_dims, self.hidden_dims, self.n_output_dims = in_dims, hidden_dims, out_dims
        self.num_layers = num_layers
        self.skip_connections = list(skip_connections)
        if num_layers == 1:
            widths = [(in_dims, out_dims)]
        else:
            widths = []
            for i in range(num_layers - 1):
                if i == 0:
                    widths.append((in_dims, hidden_dims))
                elif i in self.skip_connections:
                    widths.append((in_dims + hidden_dims, hidden_dims))
                else:
                    widths.append((hidden_dims, hidden_dims))
            widths.append((hidden_dims, out_dims))
        self.layers = nn.ModuleList([nn.Linear(a, b) for a, b in widths])

    def forward(self, x: Tensor, out_act: int = _ops.ACT_NONE) -> Tensor:
        skips = [i for i in self.skip_connections if 0 < i < len(self.layers)]
        if 0 in self.skip_connections:
            x = torch.cat([x, x], -1)             # mlp.py:42-43 with i == 0 (never used by the model)
        if len(skips) > 1:
            raise NotImplementedError("more than one skip connection")
        return _ops.mlp_chain(x, [l.weight for l in self.layers], [l.bias for l in self.layers], out_act,
                              skips[0] if skips else -1)


def run_sequential(seq: nn.Sequential, x: Tensor, out_act: int = _ops.ACT_NONE) -> Tensor:
    """Evaluate an ``nn.Sequential(Linear, ReLU, Linear, ..., [Sigmoid])`` container (the layout the
    reference uses for base/flow/shadow/dino heads, radiance_field.py:74-198) on the library's
    dense-layer kernels.  The container only holds the parameters (state-dict keys ``{0,2,4}``)."""
    linears = [m for m in seq if isinstance(m, nn.Linear)]
    if any(isinstance(m, nn.Sigmoid) for m in seq):
        out_act = _ops.ACT_SIGMOID
    return _ops.mlp_chain(x, [l.weight for l in linears], [l.bias for l in linears], out_act)
