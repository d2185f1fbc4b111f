import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import gpu_cases as G, helpers as H
import mm_distillnet_b200 as mmd
from oracle import mmd_oracle as O

def case(tag, in_grad, gscale, mta_like, default_init):
    C, CC, n_cells, B, s3 = 112, [48, 120, 352], 2, 2, 32
    gen = torch.Generator().manual_seed(0)
    torch.manual_seed(0)
    stack = mmd.BiFPNStack(*[mmd.BiFPN(C, CC, first_time=(i == 0)) for i in range(n_cells)])
    if not default_init:
        with torch.no_grad():
            for k, p in stack.named_parameters():
                if ".bn." in k or k.endswith(".1.weight") or k.endswith(".1.bias"):
                    p.copy_(torch.rand(p.shape, generator=gen) + 0.5 if k.endswith("weight") else torch.randn(p.shape, generator=gen) * 0.1)
    params = {k: v.clone() for k, v in stack.state_dict().items()}
    xs = [torch.randn(B, c, s3 >> i, s3 >> i, generator=gen) for i, c in enumerate(CC)]
    leaf = {k: (v.double().requires_grad_(True) if v.is_floating_point() and "running" not in k else (v.double() if v.is_floating_point() else v)) for k, v in params.items()}
    xr = [x.double().requires_grad_(True) for x in xs]
    tr = O.bifpn_stack(tuple(xr), leaf, n_cells, training=True)
    if mta_like:
        gouts = [(t.detach() * torch.randn(B, 1, t.shape[2], t.shape[3], generator=gen).double()).float() * gscale for t in tr]
    else:
        gouts = [torch.randn(t.shape, generator=gen) * gscale for t in tr]
    sum((t * g.double()).sum() for t, g in zip(tr, gouts)).backward()
    stack = stack.to(G.DEV).train()
    xd = [x.to(G.DEV).requires_grad_(in_grad) for x in xs]
    out = stack(tuple(xd))
    sum((t * g.to(G.DEV)).sum() for t, g in zip(out, gouts)).backward()
    res = []
    for k, p in stack.named_parameters():
        if k.endswith("conv.bias"):
            continue
        g = leaf[k].grad
        if g.abs().max() > 0:
            res.append((H.rel_l2(p.grad.cpu(), g), k, p.grad.norm().item(), g.norm().item()))
    res.sort(reverse=True)
    print(tag, " | ".join("%s %.2e (ours %.2e ref %.2e)" % (k, e, a, b) for e, k, a, b in res[:4]), flush=True)

case("rand g, in_grad, rnd init   ", True, 1.0, False, False)
case("rand g, NO in_grad, rnd init", False, 1.0, False, False)
case("rand g*1e-9, in_grad        ", True, 1e-9, False, False)
case("mta-like g, in_grad, rnd    ", True, 1.0, True, False)
case("mta-like g, in_grad, default", True, 1.0, True, True)
case("rand g, in_grad, default    ", True, 1.0, False, True)
