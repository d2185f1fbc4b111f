"""Device-timed FULL distillation step behind the backbone (BASELINE cfg 3 minus the EfficientNet-B2 backbones) at the D2 size:
student BiFPN stack + Regressor + Classifier forward/backward (train), 3 frozen teachers' stacks + heads forward (eval),
pseudo-labels from the teachers' predictions on the device, detection loss on them, 3 per-teacher MTA calls, loss combination
(src/optimization/traditional.py:171-182), backward, Adam — ModelWithNMSLoss.forward (train_methods.py:425-516) through the
drop-ins, eager and as ONE captured CUDA graph.  Synthetic backbone features; the teachers' classifier headers get a negative
bias and the confidence threshold is set to the quantile at which ~0.05 % of the anchors fire per teacher and sample (a trained detector's sparsity).
    python tools/full_step_bench.py [--batch 16] [--steps 20]"""
import argparse
import json
import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mm_distillnet_b200 as mmd   # noqa: E402
from mm_distillnet_b200 import _lib   # noqa: E402

C, CC, S3, N_CELLS, A, K, L = 112, [48, 120, 352], 96, 5, 9, 20, 3


def d2_anchors(size):
    import numpy as np
    ys = []
    for lvl in range(3, 8):
        stride = 2 ** lvl
        for sc in (2 ** 0, 2 ** (1.0 / 3.0), 2 ** (2.0 / 3.0)):
            for ra in ((1.0, 1.0), (1.4, 0.7), (0.7, 1.4)):
                hx, hy = 4.0 * stride * sc * ra[0] / 2.0, 4.0 * stride * sc * ra[1] / 2.0
                x = np.arange(stride / 2, size, stride)
                xv, yv = np.meshgrid(x, x)
                ys.append((lvl, np.stack((yv.reshape(-1) - hy, xv.reshape(-1) - hx, yv.reshape(-1) + hy, xv.reshape(-1) + hx), axis=1)))
    per_level = [np.stack([b for l, b in ys if l == lvl], axis=1).reshape(-1, 4) for lvl in range(3, 8)]
    return torch.from_numpy(np.concatenate(per_level, axis=0).astype(np.float32)).unsqueeze(0)


class Det(nn.Module):
    def __init__(self, anchors, seed):
        super().__init__()
        torch.manual_seed(seed)
        self.bifpn = mmd.BiFPNStack(*[mmd.BiFPN(C, CC, first_time=(i == 0)) for i in range(N_CELLS)])
        self.regressor, self.classifier = mmd.Regressor(C, A, L), mmd.Classifier(C, A, K, L)
        with torch.no_grad():
            self.classifier.header.pointwise_conv.conv.bias.fill_(-4.6)      # prior probability 0.01
            self.classifier.header.pointwise_conv.conv.weight.normal_(0.0, 0.25)   # a wide score distribution (no bf16 ties)
        self.anchors = anchors

    def forward(self, feats_in):
        f = self.bifpn(tuple(feats_in))
        r, _ = self.regressor(f)
        c, _ = self.classifier(f)
        return (c, r, self.anchors), f


class Feats(list):
    device = property(lambda self: self[0].device)
    shape = property(lambda self: self[0].shape)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--lockstep", action="store_true", help="student + teachers through lockstep_detection_forward (3 run_multi "
                    "calls) instead of one module call per network")
    ap.add_argument("--fire", type=float, default=0.0003, help="fraction of anchors above the confidence threshold")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    B, dt = a.batch, torch.bfloat16
    anchors = d2_anchors(768).to(dev)
    student = Det(anchors, 0).to(dev).train()
    teachers = nn.ModuleDict({m: Det(anchors, 10 + i).to(dev).eval() for i, m in enumerate(("rgb", "thermal", "depth"))})
    for p in teachers.parameters():
        p.requires_grad_(False)
    gen = torch.Generator().manual_seed(1)
    # random-init networks in eval mode (running statistics 0 / 1) let the activations decay to ~0 through 5 cells + towers: give
    # the frozen teachers "trained" BatchNorm statistics by one calibration pass with momentum 1 (running <- batch statistics)
    for m in teachers.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.momentum = 1.0

    def feats():
        return Feats([torch.randn(B, c, S3 >> i, S3 >> i, generator=gen).to(dt).to(dev).contiguous(memory_format=torch.channels_last)
                      for i, c in enumerate(CC)])
    xs, xr, xt, xd = feats(), feats(), feats(), feats()
    teachers.train()
    with torch.no_grad():
        for m, x in zip(("rgb", "thermal", "depth"), (xr, xt, xd)):
            teachers[m](x)
    teachers.eval()
    with torch.no_grad():
        smax = torch.cat([teachers[m](x)[0][0].float().max(dim=2).values.flatten() for m, x in zip(("rgb", "thermal", "depth"), (xr, xt, xd))])
    thr = float(torch.quantile(smax.cpu()[::7], 1.0 - a.fire))
    valid = list(range(K))
    vcd = {"predictions_txt2i": {"c%d" % i: i for i in valid}, "predictions_i2txt": {i: "c%d" % i for i in valid},
           "labels_txt2i": {"c%d" % i: i for i in valid}}
    cfg = {"conf_threshold": repr(thr), "nms_threshold": "0.5", "image_size": "768", "student": "YetAnotherEfficientDet"}
    model = mmd.ModelWithNMSLoss(student, teachers, mmd.YetAnotherFocalLoss(), None, mmd.MTALoss("9", "2"), cfg, vcd)
    model.pseudo_max_rows, model.pseudo_max_labels = 1024, 2048     # random detections do not cluster like a trained detector's
    opt = torch.optim.Adam(student.parameters(), lr=1e-4, betas=(0.9, 0.999), fused=True, capturable=True)

    from mm_distillnet_b200 import pseudo as PS
    crit_main, crit_kd = model.criterion_main, model.criterion_kd

    def lockstep_model():
        # ModelWithNMSLoss.forward (train_methods.py:436-516) with the networks' forwards batched across networks
        outs = mmd.lockstep_detection_forward(student, [teachers[m] for m in ("rgb", "thermal", "depth")], xs, [xr, xt, xd])
        (cs, rs, fs), touts = outs[0], outs[1:]
        with torch.no_grad():
            labels = PS.teacher_pseudo_labels([(c, r, anchors) for c, r, _ in touts], vcd, cfg, max_rows=1024, max_labels=2048)
        model.last_pseudo_labels = labels
        rl, cl = crit_main((cs, rs, anchors), labels)
        kd = crit_kd.forward_each(fs, [f for _, _, f in touts])
        return [[rl], [cl], list(kd.unbind(0))]

    def step():
        opt.zero_grad(set_to_none=True)
        out = lockstep_model() if a.lockstep else model(xr, xt, xd, xs, None)
        loss = 1.0 * (out[0][0].mean() + out[1][0].mean()) + 0.005 * torch.stack(out[2]).sum()      # traditional.py:171-182
        loss.backward()
        opt.step()
        return loss.detach()

    def timed(fn, n):
        ts = []
        for _ in range(n):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2]

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    n0 = _lib.launch_count()
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    launches = (_lib.launch_count() - n0) // 3
    model.last_pseudo_labels.check_overflow()
    counts = model.last_pseudo_labels.counts[:-1].tolist()
    with torch.cuda.stream(side):
        ms_eager = timed(step, a.steps)
    line = {"what": "full distillation step behind the backbone (student stack + heads fwd/bwd, 3 teachers' stacks + heads fwd, "
                    "device pseudo-labels, detection loss, 3 MTA calls, backward, Adam)", "batch": B, "dtype": "bf16", "lockstep": bool(a.lockstep),
            "ms_eager": round(ms_eager, 3), "samples_per_s_eager": round(B / ms_eager * 1e3, 1), "library_launches_per_step": launches,
            "labels_per_sample": counts[:8], "conf_threshold": round(thr, 5)}
    try:
        torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
    except AttributeError:
        pass
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            step()
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        ms_graph = timed(g.replay, a.steps)
        line.update(ms_graph=round(ms_graph, 3), samples_per_s_graph=round(B / ms_graph * 1e3, 1))
    except Exception as e:      # noqa: BLE001  (a timing tool: report, do not hide)
        line["graph_error"] = repr(e)[:300]
    print(json.dumps(line))


if __name__ == "__main__":
    main()
