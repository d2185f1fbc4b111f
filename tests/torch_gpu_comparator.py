#!/usr/bin/env python
"""Same-box PyTorch comparator: the plain-PyTorch restatement of the reference path (oracle/, test
infrastructure) executed by PyTorch's own CUDA kernels (cuDNN / ATen) on the same B200, same step as bench.py:
student 5-cell stack fwd+bwd (train BN) + 3 teacher stacks fwd (eval) + 3 MTA calls, B=16.

    python tests/torch_gpu_comparator.py [--batch 16] [--steps 10]

Not a pytest module (no test_ prefix) and not part of the product: it exists to put a PyTorch-on-GPU number next to
the CUDA path's.  fp32 (TF32 off) and bf16 (parameters and activations cast to bf16), NCHW and channels_last.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from oracle import mmd_oracle as O  # noqa: E402

CC, C, N_CELLS, N_T, W_KD = [48, 120, 352], 112, 5, 3, 0.005


def step(sp, tps, xs, xts):
    leaf = {k: (v.detach().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sp.items()}
    xin = [x.detach().requires_grad_(True) for x in xs]
    fs = O.bifpn_stack(tuple(xin), leaf, N_CELLS, training=True, stats_out={})
    kd = []
    for tp, xt in zip(tps, xts):
        with torch.no_grad():
            ft = O.bifpn_stack(tuple(xt), tp, N_CELLS, training=False)
        kd.append(O.mta_loss([f.float() for f in fs], [f.float() for f in ft]))
    (W_KD * torch.stack(kd).sum()).backward()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    out = []
    for dt, name in ((torch.float32, "f32"), (torch.bfloat16, "bf16")):
        for cl in (False, True):
            gen = torch.Generator().manual_seed(0)
            cast = lambda d: {k: (v.to(dev).to(dt) if v.is_floating_point() else v.to(dev)) for k, v in d.items()}
            sp = cast(O.synth_stack_params(C, CC, N_CELLS, 0))
            tps = [cast(O.synth_stack_params(C, CC, N_CELLS, 1 + k)) for k in range(N_T)]
            mk = lambda: [(lambda t: t.contiguous(memory_format=torch.channels_last) if cl else t)(
                torch.randn(a.batch, c, 96 >> i, 96 >> i, generator=gen).to(dev).to(dt)) for i, c in enumerate(CC)]
            xs, xts = mk(), [mk() for _ in range(N_T)]
            for _ in range(3):
                step(sp, tps, xs, xts)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.steps):
                step(sp, tps, xs, xts)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.steps
            r = {"dtype": name, "channels_last": cl, "ms_per_step": ms, "samples_per_s": a.batch / (ms * 1e-3)}
            out.append(r)
            print(r, flush=True)
    print(json.dumps({"workload": "bench.py step executed by PyTorch eager (oracle port) on cuda:0, B=%d" % a.batch, "rows": out}))


if __name__ == "__main__":
    main()
