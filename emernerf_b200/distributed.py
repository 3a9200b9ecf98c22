"""Ray-sharded data parallelism for the hot path (the reference itself is single-GPU, F8).

Rays are independent: every rank renders its own batch against a full replica of tables and MLPs, and the only
exchange is of parameter gradients before each optimizer step.  With :class:`emernerf_b200.optim.FusedAdam` every
gradient of a param group already lives in ONE flat buffer, so the exchange is one collective per group instead of
one per tensor (round 1 issued ~20 per step, the 122 MB table among them), in one of two forms:

``allreduce``  NCCL all-reduce (AVG) of the flat gradient, then every rank runs the same full Adam step;
``sharded``    reduce-scatter (AVG) of the flat gradient -> Adam on THIS rank's 1/world slice of the flat parameter
               space -> all-gather of the updated flat parameters.  Same bytes on the links as an all-reduce (it is its
               two halves), but the optimizer's HBM traffic (32 B per parameter, 0.15+ ms for the 203 MB of grids)
               drops by the world size.  Needs ``FusedAdam(flatten_params=True)``.

    dp = DataParallel([field_opt, prop_opt], mode="sharded")
    loss.backward(); dp.step(field_opt)          # reduce -> Adam -> (gather)

All collectives run on the current stream (they capture into the step's CUDA graph).  ``gloo`` (CPU host-logic tests)
has neither AVG nor reduce-scatter: sums + a division, and an all-reduce sliced locally, stand in.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from . import _ops
from .optim import ALIGN, FusedAdam


class DataParallel:
    def __init__(self, optimizers: Sequence[FusedAdam], mode: str = "sharded", group=None):
        if mode not in ("allreduce", "sharded"):
            raise ValueError(f"DataParallel: unknown mode {mode}")
        self.group = group
        self.on = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.world = dist.get_world_size(group) if self.on else 1
        self.rank = dist.get_rank(group) if self.on else 0
        self.native = self.on and dist.get_backend(group) == "nccl"
        self.optimizers = list(optimizers)
        self.mode = mode
        self._deferred: List[FusedAdam] = []
        self._comm = None
        if self.on and mode == "sharded":
            for opt in self.optimizers:
                for g in opt._groups:
                    if g.flat_params is None:
                        raise ValueError("DataParallel(mode='sharded') needs FusedAdam(flatten_params=True)")
                    if g.total % (ALIGN * self.world):
                        raise ValueError(f"flat size {g.total} does not split into {self.world} aligned shards")
        if self.on:
            # replicas start identical: rank 0's parameters win
            for opt in self.optimizers:
                for g in opt._groups:
                    for p in ([g.flat_params] if g.flat_params is not None else g.params):
                        dist.broadcast(p.data, src=0, group=group)

    # ------------------------------------------------------------------ pieces
    def _bounds(self, total: int):
        per = total // self.world
        return self.rank * per, (self.rank + 1) * per

    def reduce(self, opt: FusedAdam) -> None:
        """Average ``opt``'s flat gradients over the ranks (all of it, or -- sharded -- this rank's slice in place)."""
        if not self.on:
            return
        for gi, g in enumerate(opt._groups):
            if gi == len(opt._groups) - 1:
                # the LAST group holds the MLP weights, whose gradients may still be running on a side stream
                # (EMER_WGRAD_STREAM=1): the table groups in front of it are reduced meanwhile
                _ops.join_side_streams()
            flat = g.grad
            if self.mode == "allreduce":
                if self.native:
                    dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group)
                else:
                    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
                    flat.div_(self.world)
            else:
                lo, hi = self._bounds(g.total)
                if self.native:
                    dist.reduce_scatter_tensor(flat[lo:hi], flat, op=dist.ReduceOp.AVG, group=self.group)
                else:
                    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
                    flat[lo:hi].div_(self.world)
        opt.shard = (self.rank, self.world) if self.mode == "sharded" else None

    def gather(self, opt: FusedAdam) -> None:
        """Sharded mode: every rank receives the slices the others updated."""
        if not self.on or self.mode != "sharded":
            return
        for g in opt._groups:
            lo, hi = self._bounds(g.total)
            if self.native:
                dist.all_gather_into_tensor(g.flat_params, g.flat_params[lo:hi], group=self.group)
            else:
                parts = [torch.empty_like(g.flat_params[lo:hi]) for _ in range(self.world)]
                dist.all_gather(parts, g.flat_params[lo:hi].contiguous(), group=self.group)
                g.flat_params.copy_(torch.cat(parts))

    def step(self, opt: FusedAdam, defer_gather: bool = False) -> None:
        """reduce -> Adam -> gather: what replaces ``optimizer.step()`` after ``backward()``.

        ``defer_gather`` (sharded mode, for the FIELD's optimizer): the all-gather of the updated parameters is not
        issued here but by :meth:`start_deferred` at the beginning of the next step, on a side stream -- the proposal
        sampling that opens a step reads only the proposal networks, so the gather runs beside it and
        ``RadianceField.forward`` joins it (``_ops.join_before_field``)."""
        self.reduce(opt)
        opt.step()
        if defer_gather and self.on and self.mode == "sharded":
            self._deferred.append(opt)
        else:
            self.gather(opt)

    def start_deferred(self) -> None:
        """Issue the deferred all-gathers on the communication stream (call first thing in a step)."""
        if not self._deferred:
            return
        dev = self._deferred[0]._groups[0].grad.device
        if dev.type != "cuda":                     # (gloo host-logic tests: no streams)
            for opt in self._deferred:
                self.gather(opt)
            self._deferred = []
            return
        if self._comm is None:
            self._comm = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)
        self._comm.wait_stream(main)
        with torch.cuda.stream(self._comm):
            for opt in self._deferred:
                self.gather(opt)
        self._deferred = []
        _ops._BEFORE_FIELD.append(self._comm)

    def bytes_per_step(self, opt: FusedAdam) -> int:
        """Bytes each rank sends per step for ``opt`` (ring algorithms: 2 (w-1)/w of the flat size in both modes)."""
        if not self.on:
            return 0
        return int(sum(2 * (self.world - 1) / self.world * g.total * 4 for g in opt._groups))
