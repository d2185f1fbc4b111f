import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import gpu_cases as G, helpers as H
import mm_distillnet_b200 as mmd
from oracle import mmd_oracle as O

def distill_case(n_teachers, use_streams, flat, tag):
    C, CC = 112, [48, 120, 352]
    torch.manual_seed(0)
    student = mmd.BiFPNStack(*[mmd.BiFPN(C, CC, first_time=(i == 0)) for i in range(2)])
    teachers = []
    for k in range(n_teachers):
        torch.manual_seed(10 + k)
        teachers.append(mmd.BiFPNStack(*[mmd.BiFPN(C, CC, first_time=(i == 0)) for i in range(2)]))
    sp = {k: v.clone() for k, v in student.state_dict().items()}
    tps = [{k: v.clone() for k, v in t.state_dict().items()} for t in teachers]
    gen = torch.Generator().manual_seed(3)
    xs = [torch.randn(2, c, 32 >> i, 32 >> i, generator=gen) for i, c in enumerate(CC)]
    xts = [[torch.randn(2, c, 32 >> i, 32 >> i, generator=gen) * 2 for i, c in enumerate(CC)] for _ in range(n_teachers)]
    leaf = {k: (v.double().requires_grad_(True) if v.is_floating_point() and "running" not in k else (v.double() if v.is_floating_point() else v)) for k, v in sp.items()}
    fs = O.bifpn_stack(tuple(x.double() for x in xs), leaf, 2, training=True)
    kd_ref = []
    for tp, xt in zip(tps, xts):
        with torch.no_grad():
            ft = O.bifpn_stack(tuple(x.double() for x in xt), {k: (v.double() if v.is_floating_point() else v) for k, v in tp.items()}, 2, training=False)
        kd_ref.append(O.mta_loss(fs, ft))
    kd_ref = torch.stack(kd_ref)
    (0.005 * kd_ref.sum()).backward()
    student = student.to(G.DEV).train()
    teachers = [t.to(G.DEV) for t in teachers]
    step = mmd.DistillStep(student, teachers, mmd.MTALoss(), w_kd=0.005)
    if not use_streams:
        step.streams = [torch.cuda.current_stream(step.device) for _ in teachers]
    if not flat:
        student._runner.grad_sink = None
    kd = step([x.pin_memory() for x in xs], [[x.pin_memory() for x in xt] for xt in xts])
    torch.cuda.synchronize()
    worst, wn = 0.0, ""
    for k, p in student.named_parameters():
        if k.endswith("conv.bias"):
            continue
        g = leaf[k].grad
        if g.abs().max() > 0:
            e = H.rel_l2(p.grad.cpu(), g)
            if e > worst:
                worst, wn = e, "%s ours %.3e ref %.3e" % (k, p.grad.norm().item(), g.norm().item())
    print(tag, "kd err", float((kd.cpu().double() - kd_ref.detach()).abs().max()), "worst", worst, wn, flush=True)

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("c", "all"):
    distill_case(1, False, False, "1t nostream noflat")
    distill_case(1, False, True, "1t nostream flat  ")
    distill_case(1, True, True, "1t stream   flat  ")
    distill_case(2, False, False, "2t nostream noflat")
    distill_case(2, False, True, "2t nostream flat  ")
    distill_case(2, True, True, "2t stream   flat  ")
if which in ("a", "all"):
    def show(tag, m):
        print(tag, {k: round(v, 9) for k, v in m.items() if k.startswith("grad_in") or k.startswith("train_p")}, flush=True)
    show("first_s32 (fresh)", G.random_stack_case(1, True, 2, 32))
    show("first_s32 (again)", G.random_stack_case(1, True, 2, 32))
    show("2cells_s48 B1", G.random_stack_case(2, True, 1, 48))
    show("first_s32 (after)", G.random_stack_case(1, True, 2, 32))
    show("cell_s32", G.random_stack_case(1, False, 2, 32))
    show("first_s32 seed5", G.random_stack_case(1, True, 2, 32, seed=5))
