#!/usr/bin/env python
"""cProfile of the host side of the bench step (finds Python / driver overheads that hide the kernels)."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import mm_distillnet_b200 as mmd  # noqa: E402

dev = torch.device("cuda:0")
C, CC = 112, [48, 120, 352]
B = 16
dt = torch.bfloat16
torch.manual_seed(0)
student = mmd.BiFPNStack(*[mmd.BiFPN(C, CC, first_time=(i == 0)) for i in range(5)]).to(dev).train()
teachers = [mmd.BiFPNStack(*[mmd.BiFPN(C, CC, first_time=(i == 0)) for i in range(5)]).to(dev).eval() for _ in range(3)]
step = mmd.DistillStep(student, teachers, mmd.MTALoss(T=9.0, p=2.0), w_kd=0.005)
mk = lambda: [torch.randn(B, c, 96 >> i, 96 >> i, device=dev).to(dt).contiguous(memory_format=torch.channels_last)
              for i, c in enumerate(CC)]
xs, xt = mk(), [mk() for _ in range(3)]
for _ in range(5):
    step(xs, xt)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    step(xs, xt)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host enqueue %.2f ms/step, total %.2f ms/step" % ((t1 - t0) * 100, (t2 - t0) * 100))
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    step(xs, xt)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
