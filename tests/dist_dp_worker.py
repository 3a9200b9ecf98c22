"""One rank of the data-parallel step test: the PRODUCT's modules, FusedAdam and emernerf_b200.distributed.DataParallel.

    python dist_dp_worker.py <device: cpu|cuda> <mode: allreduce|sharded|single> <out.pt>

cpu: gloo, the C ABI answered by tests/cabi_emulator.py;  cuda: nccl, one GPU per rank (libemer_b200.so).
Each rank renders its half of the golden "static" batch and takes ONE optimizer step; rank 0 saves its parameters.
``single`` is the one-process full-batch step they must equal."""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.dirname(HERE), HERE, os.path.join(HERE, "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch
import torch.distributed as dist

import cases
from helpers import Golden

ADAM = dict(lr=0.01, eps=1e-15, weight_decay=1e-5, betas=(0.9, 0.99))


def main():
    device, mode, out_path = sys.argv[1], sys.argv[2], sys.argv[3]
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if device == "cpu":
        import cabi_emulator

        cabi_emulator.install(types.SimpleNamespace(setattr=setattr))
        dev = torch.device("cpu")
    else:
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
    if world > 1:
        dist.init_process_group("gloo" if device == "cpu" else "nccl", rank=rank, world_size=world)

    from emernerf_b200.distributed import DataParallel
    from emernerf_b200.optim import FusedAdam
    from emernerf_b200.radiance_fields import RadianceField, build_density_field
    from emernerf_b200.radiance_fields.encodings import HashEncoder
    from emernerf_b200.radiance_fields.render_utils import render_rays
    from emernerf_b200.third_party.nerfacc_prop_net import PropNetEstimator

    ns = types.SimpleNamespace(HashEncoder=HashEncoder, RadianceField=RadianceField, build_density_field=build_density_field)
    field, props = cases.build_models(ns, "static")
    g = Golden("static")
    field.load_state_dict(g.tensors("sd/field"))
    [p.load_state_dict(g.tensors(f"sd/prop{i}")) for i, p in enumerate(props)]
    field.to(dev).train()
    props = [p.to(dev).train() for p in props]
    est = PropNetEstimator(None, None).to(dev).train()
    opt = FusedAdam(field.parameters(), flatten_params=True, **ADAM)
    defer = mode == "sharded_defer"
    dp = DataParallel([opt], mode="sharded" if mode in ("single", "sharded_defer") else mode)

    n = cases.N_RAYS // world
    sl = slice(rank * n, (rank + 1) * n)
    batch = {k: v[sl].to(dev) for k, v in g.tensors("in/pixel").items()}
    est._jitter_override = [j[sl].to(dev) for j in g.jitters("train")]
    for step in range(2):
        dp.start_deferred()
        out = render_rays(field, est, props, batch, cases.render_cfg(), proposal_requires_grad=False)
        loss = ((out["rgb"] - batch["pixels"]) ** 2).mean() + 0.01 * out["depth"].mean()
        opt.zero_grad()
        loss.backward()
        dp.step(opt, defer_gather=defer)
    dp.start_deferred()
    if device != "cpu":
        from emernerf_b200 import _ops

        _ops.join_before_field()
        torch.cuda.synchronize()
    if rank == 0:
        torch.save({k: v.detach().cpu() for k, v in field.state_dict().items()}, out_path)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
