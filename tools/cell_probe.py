#!/usr/bin/env python
"""One non-first BiFPN cell (D2 shapes) forward/backward on cuda:0 — a short, fixed launch sequence to put under ncu.

    python tools/cell_probe.py [--batch 16] [--dtype bf16] [--mode train|eval|both] [--iters 1] [--time]

With --time it prints the per-kernel-kind CUDA-event table of the library's profiler (mmd_prof_*) instead.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import mm_distillnet_b200 as mmd  # noqa: E402
from mm_distillnet_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--mode", default="both")
    ap.add_argument("--iters", type=int, default=1)
    ap.add_argument("--cells", type=int, default=1)
    ap.add_argument("--s3", type=int, default=96)
    ap.add_argument("--time", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    dt = torch.bfloat16 if a.dtype == "bf16" else torch.float32
    torch.manual_seed(0)
    stack = mmd.BiFPNStack(*[mmd.BiFPN(112, [48, 120, 352], first_time=False) for _ in range(a.cells)]).to(dev)
    xs = [torch.randn(a.batch, 112, a.s3 >> i, a.s3 >> i, device=dev).to(dt).contiguous(memory_format=torch.channels_last)
          for i in range(5)]

    def train_step():
        stack.train()
        xi = [x.detach().requires_grad_(True) for x in xs]
        outs = stack(tuple(xi))
        torch.autograd.backward(outs, [torch.ones_like(o) for o in outs])

    def eval_step():
        stack.eval()
        with torch.no_grad():
            stack(tuple(xs))

    def run():
        if a.mode in ("train", "both"):
            train_step()
        if a.mode in ("eval", "both"):
            eval_step()

    for _ in range(2):
        run()
    torch.cuda.synchronize()
    if a.time:
        _lib.prof_enable(True)
        _lib.prof_collect()
    for _ in range(a.iters):
        run()
    torch.cuda.synchronize()
    if a.time:
        prof = _lib.prof_collect()
        _lib.prof_enable(False)
        for k, v in prof.items():
            if v["launches"]:
                print("%-12s launches %4d  avg %8.1f us  %8.1f GB/s (algorithmic)" % (
                    k, v["launches"], 1e3 * v["ms"] / v["launches"],
                    v["algo_bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else 0.0))


if __name__ == "__main__":
    main()
