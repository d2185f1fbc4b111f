import os, sys, time, traceback
sys.path.insert(0, "/root/repo")
import torch
import mm_distillnet_b200 as mmd
from mm_distillnet_b200 import bifpn
dev = torch.device("cuda:0")
C, CC = 112, [48, 120, 352]
B = 16
dt = torch.bfloat16
torch.manual_seed(0)
student = mmd.BiFPNStack(*[mmd.BiFPN(C, CC, first_time=(i == 0)) for i in range(5)]).to(dev).train()
teachers = [mmd.BiFPNStack(*[mmd.BiFPN(C, CC, first_time=(i == 0)) for i in range(5)]).to(dev).eval() for _ in range(3)]
step = mmd.DistillStep(student, teachers, mmd.MTALoss(T=9.0, p=2.0), w_kd=0.005)
mk = lambda: [torch.randn(B, c, 96 >> i, 96 >> i, device=dev).to(dt).contiguous(memory_format=torch.channels_last) for i, c in enumerate(CC)]
xs, xt = mk(), [mk() for _ in range(3)]
xs = [x.requires_grad_(True) for x in xs]
for _ in range(5):
    step(xs, xt)
torch.cuda.synchronize()
orig = bifpn._Lease.__init__
cnt = {"hit": 0, "miss": 0, "t": 0.0}
def init(self, pool, nbytes, device):
    t0 = time.perf_counter()
    if pool: cnt["hit"] += 1
    else: cnt["miss"] += 1
    orig(self, pool, nbytes, device)
    cnt["t"] += time.perf_counter() - t0
bifpn._Lease.__init__ = init
for k in range(5):
    t0 = time.perf_counter()
    for x in xs: x.grad = None
    step(xs, xt)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("step", k, "enqueue %.2f ms total %.2f ms" % ((t1-t0)*1e3, (t2-t0)*1e3), cnt)
print(torch.cuda.memory_stats()["num_alloc_retries"], torch.cuda.memory_stats()["num_device_alloc"], torch.cuda.memory_stats()["num_device_free"])
