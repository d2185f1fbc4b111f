"""Three eager distillation steps (student + 3 teachers in lockstep, MTA, backward) at batch B: the workload the ncu
captures of profiles/ are taken on.  With --chain the persistent small-level chains are switched on and the library
prints %globaltimer deltas for every step of every chain launch (MMD_CHAIN_DEBUG=1).
    python tools/step_probe.py [B] [--chain]"""
import os
import sys

if "--chain" in sys.argv:
    sys.argv.remove("--chain")
    os.environ.setdefault("MMD_CHAIN", "1")
    os.environ.setdefault("MMD_CHAIN_DEBUG", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import mm_distillnet_b200 as mmd  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda", 0)
torch.manual_seed(0)
student = mmd.BiFPNStack(*[mmd.BiFPN(112, bench.CC, first_time=(i == 0)) for i in range(5)]).to(dev).train()
teachers = [mmd.BiFPNStack(*[mmd.BiFPN(112, bench.CC, first_time=(i == 0)) for i in range(5)]).to(dev).eval() for _ in range(3)]
step = mmd.DistillStep(student, teachers, mmd.MTALoss(), w_kd=0.005)
gen = torch.Generator().manual_seed(1)
mk = lambda: [x.to(dev).contiguous(memory_format=torch.channels_last) for x in bench.synth_inputs(B, gen, torch.bfloat16)]
xs, xt = mk(), [mk() for _ in range(3)]
for it in range(3):
    print("---- step", it, file=sys.stderr, flush=True)
    step(xs, xt)
    torch.cuda.synchronize()
