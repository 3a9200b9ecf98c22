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
        skip_in = x
        last = len(self.layers) - 1
        for i, layer in enumerate(self.layers):
            if i in self.skip_connections:
                x = torch.cat([x, skip_in], -1)
            x = _ops.linear(x, layer.weight, layer.bias, _ops.ACT_RELU if i < last else out_act)
        return x


def run_sequential(seq: nn.Sequential, x: Tensor, out_act: int = _ops.ACT_NONE) -> Tensor:
    """Evaluate an ``nn.Sequential(Linear, ReLU, Linear, ..., [Sigmoid])`` container (the layout the
    reference uses for base/flow/shadow/dino heads, radiance_field.py:74-198) on the library's
    dense-layer kernels.  The container only holds the parameters (state-dict keys ``{0,2,4}``)."""
    linears = [m for m in seq if isinstance(m, nn.Linear)]
    if any(isinstance(m, nn.Sigmoid) for m in seq):
        out_act = _ops.ACT_SIGMOID
    for i, lin in enumerate(linears):
        x = _ops.linear(x, lin.weight, lin.bias, _ops.ACT_RELU if i < len(linears) - 1 else out_act)
    return x
