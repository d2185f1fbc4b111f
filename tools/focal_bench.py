"""Device-timed YetAnotherFocalLoss forward+backward at the D2 size (110 484 anchors, 20 classes) against the
HBM roofline.  Algorithmic bytes: forward reads B*N*(K+4) elements once, backward reads them once and writes as many.
    python tools/focal_bench.py [--batch 16] [--boxes 8] [--f32]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mm_distillnet_b200 as mmd   # noqa: E402
from mm_distillnet_b200 import _lib   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--boxes", type=int, default=8)
    ap.add_argument("--classes", type=int, default=20)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--f32", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    dt = torch.float32 if a.f32 else torch.bfloat16
    es = 4 if a.f32 else 2
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    N, K, B = 110484, a.classes, a.batch
    gen = torch.Generator().manual_seed(0)
    c = torch.rand(B, N, K, generator=gen).pow(3.0).to(dev).to(dt).requires_grad_(True)
    r = (0.5 * torch.randn(B, N, 4, generator=gen)).to(dev).to(dt).requires_grad_(True)
    xy = torch.rand(N, 2, generator=gen) * 700
    wh = 16 + torch.rand(N, 2, generator=gen) * 200
    anchors = torch.cat([xy, xy + wh], dim=1).unsqueeze(0).to(dev)
    ann = []
    for b in range(B):
        p = torch.rand(a.boxes, 2, generator=gen) * 500
        s = 40 + torch.rand(a.boxes, 2, generator=gen) * 250
        ann.append(torch.cat([p, p + s, torch.randint(0, K, (a.boxes, 1), generator=gen).float()], dim=1).numpy().astype(np.float32))
    crit = mmd.YetAnotherFocalLoss()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step():
        c.grad = None
        r.grad = None
        rl, cl = crit((c, r, anchors), ann)
        (rl + cl).sum().backward()

    n0 = _lib.launch_count()
    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    per_step = (_lib.launch_count() - n0) // a.warmup
    ts = []
    for _ in range(a.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    algo = 3 * B * N * (K + 4) * es
    _lib.lib().mmd_prof_enable(1)
    step()
    import ctypes as C
    nk = _lib.lib().mmd_prof_num_kinds()
    t, cnt, by = (C.c_double * nk)(), (C.c_longlong * nk)(), (C.c_double * nk)()
    _lib.lib().mmd_prof_collect(t, cnt, by)
    _lib.lib().mmd_prof_enable(0)
    kern = {}
    for i in range(nk):
        if cnt[i]:
            nm = _lib.lib().mmd_prof_kind_name(i).decode()
            kern[nm] = {"launches": int(cnt[i]), "ms": round(t[i], 4), "GBps": round(by[i] / t[i] / 1e6, 1), "frac": round(by[i] / t[i] / 1e6 / peak, 4)}
    print(json.dumps({"workload": "YetAnotherFocalLoss forward+backward, N=%d anchors, K=%d classes, B=%d, %d boxes/sample" % (N, K, B, a.boxes),
                      "dtype": "f32" if a.f32 else "bf16", "ms_per_step": round(ms, 4), "ms_min": round(ts[0], 4),
                      "samples_per_s": round(B / ms * 1e3, 1), "gpu_launches_per_step": per_step,
                      "roofline": {"bound": "hbm", "algo_bytes": algo, "achieved": round(algo / ms / 1e6, 1), "peak": peak, "unit": "GB/s",
                                   "frac": round(algo / ms / 1e6 / peak, 4)},
                      "kernels": kern,
                      "timing": "median of %d eager steps incl. the host-side label padding + H2D copy, CUDA events, L2 flushed" % a.steps}))


if __name__ == "__main__":
    main()
