"""One rank of the world-size-2 gloo test (launched by tests/test_distributed_cpu.py)."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.dirname(HERE), HERE, os.path.join(HERE, "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch
import torch.distributed as dist

import cases
from helpers import Golden


def loss_and_grads(g: Golden, sl: slice):
    from test_oracle_golden import _specs
    from oracle import hotpath

    fs, ps = _specs("static")
    fsd = g.tensors("sd/field")
    psd = [g.tensors(f"sd/prop{i}") for i in range(2)]
    keys = [k for k in fsd if k.endswith(("weight", "bias", "params"))]
    for k in keys:
        fsd[k].requires_grad_(True)
    batch = {k: v[sl] for k, v in g.tensors("in/pixel").items()}
    out, _ = hotpath.render_rays(fsd, fs, psd, ps, batch, num_samples=cases.NUM_SAMPLES,
                                 prop_samples=cases.PROP_SAMPLES, near_plane=cases.NEAR, far_plane=cases.FAR,
                                 training=False)
    loss = ((out["rgb"] - batch["pixels"]) ** 2).mean()
    grads = torch.autograd.grad(loss, [fsd[k] for k in keys], allow_unused=True)
    return keys, [torch.zeros_like(fsd[k]) if gr is None else gr for k, gr in zip(keys, grads)]


if __name__ == "__main__":
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = Golden("static")
    n = cases.N_RAYS // world
    keys, grads = loss_and_grads(g, slice(rank * n, (rank + 1) * n))
    for gr in grads:                         # the recipe of bench.py:Trainer.allreduce
        dist.all_reduce(gr, op=dist.ReduceOp.SUM)
        gr /= world
    if rank == 0:
        torch.save({k: gr for k, gr in zip(keys, grads)}, sys.argv[1])
    dist.barrier()
    dist.destroy_process_group()
