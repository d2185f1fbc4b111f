"""(Lives under tests/ because it takes its synthetic detector outputs and the host-side comparison from the oracle.)
Device-timed pseudo-label generation at the D2 size (110 484 anchors, 20 classes, 3 teachers) — alone and
followed by the detection loss on the device-made labels — against the HBM roofline, next to the restated reference path
(the oracle: per-sample Python + torch CPU NMS, with the device->host copies the reference pays) on the same inputs.
Algorithmic bytes of the generation: the T*B*N*(K+4) prediction elements read once.
    python tests/pseudo_bench.py [--batch 16] [--f32] [--objects 7] [--cpu]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mm_distillnet_b200 as mmd   # noqa: E402
from mm_distillnet_b200 import _lib   # noqa: E402
from mm_distillnet_b200 import pseudo as PS   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--teachers", type=int, default=3)
    ap.add_argument("--objects", type=int, default=7)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--f32", action="store_true")
    ap.add_argument("--cpu", action="store_true", help="also time the restated reference path (tests' oracle) on the host")
    a = ap.parse_args()
    from oracle import mmd_oracle as O          # bench tool: inputs come from the oracle's synthetic detector outputs
    from tests import helpers as H
    dev = torch.device("cuda:0")
    dt = torch.float32 if a.f32 else torch.bfloat16
    es = 4 if a.f32 else 2
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    size, K, B, T = 768, 20, a.batch, a.teachers
    anchors = H.efficientdet_anchors(size)
    N = anchors.shape[1]
    base = [O.synth_teacher_logits(4, anchors, K, 300 + 10 * t, n_objects=a.objects, size=size) for t in range(T)]
    rep = (B + 3) // 4
    logits = [(c.repeat(rep, 1, 1)[:B].contiguous(), r.repeat(rep, 1, 1)[:B].contiguous()) for c, r in base]
    preds = [(c.to(dev).to(dt), r.to(dev).to(dt), anchors.to(dev)) for c, r in logits]
    vcd, cfg = H.pseudo_valid_classes_dict(), H.pseudo_config(size)
    crit = mmd.YetAnotherFocalLoss()
    cs = torch.rand(B, N, K).pow(3.0).to(dev).to(dt).requires_grad_(True)
    rs = (0.5 * torch.randn(B, N, 4)).to(dev).to(dt).requires_grad_(True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def gen():
        return PS.teacher_pseudo_labels(preds, vcd, cfg)

    def gen_and_loss():
        cs.grad = None
        rs.grad = None
        rl, cl = crit((cs, rs, anchors.to(dev)), gen())
        (rl + cl).sum().backward()

    def timed(fn):
        n0 = _lib.launch_count()
        for _ in range(a.warmup):
            fn()
        torch.cuda.synchronize()
        per = (_lib.launch_count() - n0) // a.warmup
        ts = []
        for _ in range(a.steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2], per

    ms_gen, l_gen = timed(gen)
    ms_all, l_all = timed(gen_and_loss)
    out = gen()
    out.check_overflow()
    _lib.prof_enable(True)
    gen()
    kern = _lib.prof_collect()
    _lib.prof_enable(False)
    algo = T * B * N * (K + 4) * es
    line = {"what": "pseudo-label generation (%d teachers) on device" % T, "batch": B, "dtype": "bf16" if not a.f32 else "f32",
            "anchors": N, "classes": K, "labels_per_sample": [int(x) for x in out.counts[:-1].cpu().tolist()][:8],
            "rows_per_teacher_sample": [int(x) for x in out.teacher_counts.cpu().flatten().tolist()][:8],
            "ms_generate": round(ms_gen, 4), "launches_generate": l_gen, "samples_per_s_generate": round(B / ms_gen * 1e3, 1),
            "algo_bytes": algo, "gbs": round(algo / ms_gen / 1e6, 1), "frac_of_hbm_peak": round(algo / ms_gen / 1e6 / peak, 4),
            "ms_generate_plus_focal_fwd_bwd": round(ms_all, 4), "launches_generate_plus_focal": l_all,
            "kernels": {k: {"ms": round(v["ms"], 4), "launches": v["launches"],
                            "frac": round(v["algo_bytes"] / v["ms"] / 1e6 / peak, 4) if v["ms"] > 0 else None} for k, v in kern.items()},
            "data": "synthetic detector-like outputs (oracle.synth_teacher_logits), L2 flushed between steps"}
    if a.cpu:
        # the reference's path on the same inputs: predictions live on the device, its post-processing thresholds there and
        # moves every sample's candidates to the host (utils.py:217-221), NMS + integration on the host, labels go back inside
        # the detection loss.  Restated by the oracle on host tensors + the D2H of the predictions' candidates is NOT charged:
        # a lower bound for the reference.
        label_of = {i: n for n, i in enumerate(H.PSEUDO_VALID_IDS)}
        lf = [(c.to(dt).float(), r.to(dt).float()) for c, r in logits]
        t0 = time.perf_counter()
        per_t = [O.logits_to_ground_truth((c, r, anchors), H.PSEUDO_VALID_IDS, label_of, image_size=size, include_scores=True,
                                          **H.PSEUDO_CFG) for c, r in lf]
        O.merge_teacher_labels(per_t)
        line["cpu_port_ms"] = round((time.perf_counter() - t0) * 1e3, 1)
        line["cpu_port_threads"] = torch.get_num_threads()
    print(json.dumps(line))


if __name__ == "__main__":
    main()
