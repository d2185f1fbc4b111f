"""Probe (timing only): does running the frozen teachers' forward on a side stream NEXT TO the student's forward + backward
(software pipelining across steps: teachers run one batch ahead) beat the lockstep step?  Same kernels and work per
step; numerics are not checked here (the KD call reads the previous iteration's teacher features).
    python tools/pipe_probe.py [--batch 32]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mm_distillnet_b200 as mmd   # noqa: E402
from mm_distillnet_b200.bifpn import forward_multi   # noqa: E402

C, CC, S3, N_CELLS, W_KD = 112, [48, 120, 352], 96, 5, 0.005


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--mode", default="all", choices=["all", "lockstep", "pipe_bwd", "pipe_all"])
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    B, dt = a.batch, torch.bfloat16
    torch.manual_seed(0)
    student = mmd.BiFPNStack(*[mmd.BiFPN(C, CC, first_time=(i == 0)) for i in range(N_CELLS)]).to(dev).train()
    teachers = [mmd.BiFPNStack(*[mmd.BiFPN(C, CC, first_time=(i == 0)) for i in range(N_CELLS)]).to(dev).eval() for _ in range(3)]
    for t in teachers:
        for p in t.parameters():
            p.requires_grad_(False)
    crit = mmd.MTALoss(9.0, 2.0)
    gen = torch.Generator().manual_seed(1)

    def inputs(grad):
        return [torch.randn(B, c, S3 >> i, S3 >> i, generator=gen).to(dt).to(dev).contiguous(memory_format=torch.channels_last)
                .requires_grad_(grad) for i, c in enumerate(CC)]
    xs = inputs(True)
    xts = [inputs(False) for _ in range(3)]
    side = torch.cuda.Stream(device=dev)
    g = torch.full((3, 5), W_KD, device=dev)

    def teachers_fwd():
        with torch.no_grad():
            return forward_multi([(t, x) for t, x in zip(teachers, xts)])

    def lockstep():
        outs = forward_multi([(student, xs)] + [(t, x) for t, x in zip(teachers, xts)])
        kd = crit.forward_each(outs[0], [[f.detach() for f in o] for o in outs[1:]])
        torch.autograd.backward([kd], [g])

    state = {"feats": None}

    def pipelined(overlap_fwd):
        main = torch.cuda.current_stream(dev)
        if state["feats"] is None:
            state["feats"] = [[f.detach().clone() for f in o] for o in teachers_fwd()]
        if overlap_fwd:
            side.wait_stream(main)
            with torch.cuda.stream(side):
                new = teachers_fwd()
        fs = student(tuple(xs))
        kd = crit.forward_each(fs, state["feats"])
        if not overlap_fwd:
            side.wait_stream(main)
            with torch.cuda.stream(side):
                new = teachers_fwd()
        else:
            ev = torch.cuda.Event()
            ev.record(main)
            side.wait_event(ev)
        with torch.cuda.stream(side), torch.no_grad():
            for dst, src in zip(state["feats"], new):
                for d, s in zip(dst, src):
                    d.copy_(s)
        torch.autograd.backward([kd], [g])
        main.wait_stream(side)

    def measure(fn, name):
        cap = torch.cuda.Stream(device=dev)
        cap.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(cap):
            for _ in range(3):
                for x in xs:
                    x.grad = None
                fn()
        torch.cuda.current_stream(dev).wait_stream(cap)
        torch.cuda.synchronize()
        for x in xs:
            x.grad = None
        try:
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        except AttributeError:
            pass
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=cap):
            fn()
        for _ in range(5):
            graph.replay()
        torch.cuda.synchronize()
        ts = []
        for _ in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.steps):
                graph.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / a.steps)
        ts.sort()
        print(json.dumps({"mode": name, "batch": B, "ms_per_step": round(ts[len(ts) // 2], 4), "min": round(ts[0], 4),
                          "samples_per_s": round(B / ts[len(ts) // 2] * 1e3, 1)}))

    if a.mode in ("all", "lockstep"):
        measure(lockstep, "lockstep (shipped)")
    if a.mode in ("all", "pipe_bwd"):
        measure(lambda: pipelined(False), "teachers(next) on a side stream next to the student's backward")
    if a.mode in ("all", "pipe_all"):
        state["feats"] = None
        measure(lambda: pipelined(True), "teachers(next) on a side stream next to the student's forward + backward")


if __name__ == "__main__":
    main()
