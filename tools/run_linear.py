"""Launch the dense-layer kernels once at bench shapes (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from emernerf_b200 import _ops

n = 524288
for k, o, act in ((64, 64, 1), (177, 64, 1)):
    x = torch.randn(n, k, device="cuda", requires_grad=True)
    w = (torch.randn(o, k, device="cuda") / k ** 0.5).requires_grad_(True)
    b = torch.zeros(o, device="cuda", requires_grad=True)
    for _ in range(3):
        y = _ops.linear(x, w, b, act)
        y.backward(torch.ones_like(y))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); y = _ops.linear(x, w, b, act); e1.record(); torch.cuda.synchronize()
    print(k, o, "fwd ms", e0.elapsed_time(e1))

# per-phase cycle counters of CTA 0 (debug hook)
import ctypes
from emernerf_b200 import _lib
names = ["issue_cp", "wait_cp", "wait_empty", "convert", "fence_sync", "mma_issue", "wait_accum", "epilogue"]
for k, o in ((64, 64), (180, 64)):
    x = torch.randn(n, k, device="cuda"); w = torch.randn(o, k, device="cuda"); b = torch.zeros(o, device="cuda")
    dbg = torch.zeros(10, dtype=torch.int64, device="cuda")
    _lib.call("emer_debug_tc_timing", ctypes.c_void_p(dbg.data_ptr()))
    y = _ops.linear(x, w, b, 1)
    torch.cuda.synchronize()
    _lib.call("emer_debug_tc_timing", None)
    d = dbg.tolist()
    tot = sum(d[:8])
    print(f"k={k} o={o}: chunks {d[8]} tiles {d[9]}  total {tot} cyc = {tot / max(d[9],1):.0f} cyc/tile")
    print("   " + "  ".join(f"{nm}={v / max(d[9],1):.0f}" for nm, v in zip(names, d[:8])))
