import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import gpu_cases as G, helpers as H
import mm_distillnet_b200 as mmd
from oracle import mmd_oracle as O
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("a", "all"):
    for rep in range(3):
        m = G.random_stack_case(2, True, 1, 48)
        print("rep", rep, {k: (round(v, 9) if not isinstance(v, str) else v) for k, v in m.items() if k.startswith("grad_in") or k.startswith("pgrad") or k.startswith("worst")})
    m = G.random_stack_case(1, True, 2, 32)
    print("other", {k: round(v, 9) for k, v in m.items() if k.startswith("grad_in")})
    m = G.random_stack_case(2, True, 1, 48)
    print("again", {k: round(v, 9) for k, v in m.items() if k.startswith("grad_in")})
if which in ("c", "all"):
    C, CC = 112, [48, 120, 352]
    torch.manual_seed(0)
    student = mmd.BiFPNStack(*[mmd.BiFPN(C, CC, first_time=(i == 0)) for i in range(2)])
    torch.manual_seed(10)
    teacher = mmd.BiFPNStack(*[mmd.BiFPN(C, CC, first_time=(i == 0)) for i in range(2)])
    sp = {k: v.clone() for k, v in student.state_dict().items()}
    tp = {k: v.clone() for k, v in teacher.state_dict().items()}
    gen = torch.Generator().manual_seed(3)
    xs = [torch.randn(2, c, 32 >> i, 32 >> i, generator=gen) for i, c in enumerate(CC)]
    xt = [torch.randn(2, c, 32 >> i, 32 >> i, generator=gen) * 2 for i, c in enumerate(CC)]
    leaf = {k: (v.double().requires_grad_(True) if v.is_floating_point() and "running" not in k else (v.double() if v.is_floating_point() else v)) for k, v in sp.items()}
    fs = O.bifpn_stack(tuple(x.double() for x in xs), leaf, 2, training=True)
    with torch.no_grad():
        ft = O.bifpn_stack(tuple(x.double() for x in xt), {k: (v.double() if v.is_floating_point() else v) for k, v in tp.items()}, 2, training=False)
    kd_ref = O.mta_loss(fs, ft)
    gf = torch.autograd.grad((0.005 * kd_ref.sum()), fs, retain_graph=True)
    (0.005 * kd_ref.sum()).backward()
    student = student.to(G.DEV).train(); teacher = teacher.to(G.DEV).eval()
    # plain autograd path (no DistillStep)
    xd = [x.to(G.DEV) for x in xs]
    f = student(tuple(xd))
    with torch.no_grad():
        t = teacher(tuple(x.to(G.DEV) for x in xt))
    kd = mmd.MTALoss()(f, t)
    gf_ours = torch.autograd.grad(0.005 * kd.sum(), f, retain_graph=True)
    for a, b in zip(gf_ours, gf):
        print("dL/dfeat rel", H.rel_l2(a.cpu(), b), "norm", b.norm().item())
    (0.005 * kd.sum()).backward()
    print("kd", kd.tolist(), kd_ref.tolist())
    for k, p in list(student.named_parameters())[:12] + list(student.named_parameters())[-6:]:
        g = leaf[k].grad
        print("%-50s ours %.4e ref %.4e rel %.3e" % (k, p.grad.norm().item(), g.norm().item(), H.rel_l2(p.grad.cpu(), g)))
