"""Device-timed forward+backward of the detection heads at the cfg2 pyramid (768x768 input: 96..6, C=112,
9 anchors), bf16 storage, against the HBM roofline.  One JSON line per head.

Algorithmic bytes per sample and forward (same convention as the BiFPN nodes): every tower layer reads and
writes one C-channel map, the header reads one and writes K channels: 12 276 positions x (7*112 + K) elements; a training
step (forward + backward) is charged 3x that.
    python tools/heads_bench.py [--batch 16] [--classes 1] [--steps 20] [--f32]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mm_distillnet_b200 as mmd   # noqa: E402
from mm_distillnet_b200 import _lib   # noqa: E402

SIZES = (96, 48, 24, 12, 6)
POS = sum(s * s for s in SIZES)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--classes", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--f32", action="store_true")
    ap.add_argument("--eval", action="store_true", help="forward only, eval mode (a frozen teacher's heads)")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    dt = torch.float32 if a.f32 else torch.bfloat16
    es = 4 if a.f32 else 2
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    peak = float(peaks["hbm_gbs"])
    torch.manual_seed(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name, mod, K in (("regressor", mmd.Regressor(112, 9, 3), 36),
                         ("classifier", mmd.Classifier(112, 9, a.classes, 3), 9 * a.classes)):
        mod = mod.to(dev)
        mod.train(not a.eval)
        xs = [torch.randn(a.batch, 112, s, s, device=dev).to(dt).contiguous(memory_format=torch.channels_last).requires_grad_(not a.eval)
              for s in SIZES]

        def step():
            if a.eval:
                with torch.no_grad():
                    return mod(xs)
            mod.zero_grad(set_to_none=True)
            for x in xs:
                x.grad = None
            y, al = mod(xs)
            gy, ga = torch.ones_like(y), torch.ones_like(al)
            torch.autograd.backward([y, al], [gy, ga])
            return y

        n0 = _lib.launch_count()
        for _ in range(a.warmup):
            step()
        torch.cuda.synchronize()
        per_step = (_lib.launch_count() - n0) // a.warmup
        ts = []
        for _ in range(a.steps):
            flush.zero_()                                    # L2 flush between timed iterations
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        ms = ts[len(ts) // 2]
        fwd_bytes = a.batch * POS * (7 * 112 + K) * es
        algo = fwd_bytes * (1 if a.eval else 3)
        print(json.dumps({"workload": "%s %s, cfg2 pyramid 96..6, C=112, 9 anchors, K=%d, B=%d" %
                          (name, "eval forward" if a.eval else "train forward+backward", K, a.batch),
                          "dtype": "f32" if a.f32 else "bf16", "ms_per_step": round(ms, 4), "ms_min": round(ts[0], 4),
                          "samples_per_s": round(a.batch / ms * 1e3, 1), "gpu_launches_per_step": per_step,
                          "roofline": {"bound": "hbm", "algo_bytes": algo, "achieved": round(algo / ms / 1e6, 1), "peak": peak,
                                       "unit": "GB/s", "frac": round(algo / ms / 1e6 / peak, 4)},
                          "timing": "median of %d eager steps (host launch included), CUDA events, 256 MB L2 flush between steps" % a.steps}))


if __name__ == "__main__":
    main()
