#!/usr/bin/env python
"""BASELINE cfg 5 on the hot path: student-only inference (evaluate.py path: `model.eval()`, no_grad) through the 5-cell
BiFPN stack, batch sweep, samples/s on one B200.

    python tools/eval_sweep.py [--dtype bf16|f32] [--batches 1,2,4,...,256] [--iters 20]

Inputs are synthetic C3/C4/C5 of a 768x768 image, resident in HBM; time = CUDA events around `iters` forwards after
5 warm-up forwards (inputs >= 2 batches are rotated so that consecutive forwards do not reread the same lines from L2).
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import mm_distillnet_b200 as mmd  # noqa: E402

CC = [48, 120, 352]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--batches", default="1,2,4,8,16,32,64,128,256")
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    dt = torch.bfloat16 if a.dtype == "bf16" else torch.float32
    torch.manual_seed(0)
    stack = mmd.BiFPNStack(*[mmd.BiFPN(112, CC, first_time=(i == 0)) for i in range(5)]).to(dev).eval()
    rows = []
    for B in [int(b) for b in a.batches.split(",")]:
        sets = [[torch.randn(B, c, 96 >> i, 96 >> i, device=dev).to(dt).contiguous(memory_format=torch.channels_last)
                 for i, c in enumerate(CC)] for _ in range(2)]
        with torch.no_grad():
            for k in range(5):
                stack(tuple(sets[k & 1]))
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for k in range(a.iters):
                stack(tuple(sets[k & 1]))
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        rows.append({"batch": B, "ms_per_forward": ms, "samples_per_s": B / (ms * 1e-3)})
        print("B=%4d  %8.3f ms/forward  %10.1f samples/s" % (B, ms, B / (ms * 1e-3)), flush=True)
        del sets
        torch.cuda.empty_cache()
    print(json.dumps({"workload": "cfg5: student-only eval forward, 5-cell D2 BiFPN stack, %s" % a.dtype, "rows": rows}))


if __name__ == "__main__":
    main()
