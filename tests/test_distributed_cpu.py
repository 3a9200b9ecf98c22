"""World-size-2 gloo test of the multi-GPU recipe on CPU: ray sharding + gradient averaging gives the
same parameter gradient as the single-process full batch (rays are independent units; the only
collective is the gradient all-reduce).  The CPU oracle plays the model."""
import os
import socket
import subprocess
import sys

import torch

import cases
from helpers import Golden

HERE = os.path.dirname(os.path.abspath(__file__))


def test_ray_sharded_gradient_average_equals_full_batch(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "grads.pt")
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "dist_worker.py"), out], env=env))
    assert all(p.wait(timeout=240) == 0 for p in procs)
    sharded = torch.load(out)
    from dist_worker import loss_and_grads

    keys, full = loss_and_grads(Golden("static"), slice(0, cases.N_RAYS))
    for k, gr in zip(keys, full):
        denom = gr.abs().max().clamp_min(1e-12)
        assert ((sharded[k] - gr).abs().max() / denom).item() < 1e-5, k
