"""MTA loss: CUDA path (MTALoss module -> mmd_mta_fwd / mmd_mta_bwd) vs the reference's golden outputs and the
fp64 oracle.  Tolerances (fp32): |loss - ref| <= 2e-6 absolute (the loss is -ln N - 1/N + O(1e-4), so this is the
resolution of an fp32 loss); gradient rel-L2 <= 1e-4 against the fp64 oracle, in the gradient's own scale
(SURVEY.md 0.2).  bf16: gradients are STORED in bf16 (2^-9 relative rounding) -> rel-L2 <= 4e-3."""
import numpy as np
import pytest
import torch

from tests import gpu_cases as G
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,cl", [("mta_c112", False), ("mta_c112", True), ("mta_c16", True), ("mta_c16", False)])
def test_mta_golden(name, cl):
    m = G.mta_golden_case(name, channels_last=cl)
    assert m["loss_abs_single"] <= 2e-6 and m["loss_abs_multi"] <= 2e-6, m
    # the golden gradients come from the fp32 reference, which itself is ~1e-5..1e-4 from fp64 (cancellation)
    assert m["grad_single"] <= 2e-4 and m["grad_multi"] <= 2e-4, m


@pytest.mark.parametrize("nt", [1, 2, 3, 4])
def test_mta_vs_fp64_oracle(nt):
    m = G.mta_random_case(4, 112, [24, 12, 6, 3, 2], nt)
    assert m["loss_abs"] <= 2e-6, m
    assert m["grad"] <= 1e-4, m


@pytest.mark.parametrize("nt,dtype,tol", [(1, torch.float32, 1e-4), (3, torch.float32, 1e-4), (4, torch.float32, 1e-4),
                                          (3, torch.bfloat16, 4e-3)])
def test_mta_forward_each_matches_per_teacher_calls(nt, dtype, tol):
    """One batched launch set == the reference's per-teacher criterion_kd calls (train_methods.py:351-358)."""
    m = G.mta_each_case(4, 112, [24, 12, 6, 3, 2], nt, dtype=dtype)
    assert m["shape_ok"], m
    assert m["loss_abs"] <= 2e-6 and m["loss_vs_loop"] <= 1e-6, m
    assert m["grad"] <= tol, m
    # fp32: the summed gradient equals the sum autograd forms from nt separate calls up to fp32 summation order;
    # bf16: the loop rounds every call's gradient to bf16 before summing, the batched pass rounds once
    assert m["grad_vs_loop"] <= (1e-5 if dtype == torch.float32 else 8e-3), m


def test_mta_full_size_uniform_landmarks():
    """BASELINE sizes (B=16, P3..P7 of a 768^2 input), unstructured features: loss = -ln(HW) - 1/HW to 1e-3."""
    import mm_distillnet_b200 as mmd
    torch.manual_seed(0)
    sizes = [96, 48, 24, 12, 6]
    fs = [torch.randn(16, 112, s, s, device=G.DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True) for s in sizes]
    ft = [[torch.randn(16, 112, s, s, device=G.DEV).contiguous(memory_format=torch.channels_last) for s in sizes] for _ in range(3)]
    crit = mmd.MTALoss()
    l1 = crit(fs, ft[0])
    for l, s in zip(l1.tolist(), sizes):
        assert abs(l - (-np.log(s * s) - 1.0 / (s * s))) < 1e-3
    # linearity of the backward in grad_output and teacher-order invariance of the product branch
    l3 = crit(fs, ft)
    l3p = crit(fs, [ft[2], ft[0], ft[1]])
    assert torch.allclose(l3, l3p, rtol=0, atol=1e-6)
    g1 = torch.autograd.grad((l3 * 0.005).sum(), fs, retain_graph=True)
    g2 = torch.autograd.grad((l3 * 0.010).sum(), fs)
    for a, b in zip(g1, g2):
        assert H.rel_l2(2 * a, b) < 1e-6
        assert a.shape == fs[0].shape[:2] + a.shape[2:]


def test_mta_bf16_and_layouts():
    m = G.mta_random_case(4, 112, [24, 12, 6], 2, dtype=torch.bfloat16)
    assert m["loss_abs"] <= 2e-6 and m["grad"] <= 4e-3, m
    m = G.mta_random_case(3, 112, [12, 6], 1, channels_last=False)
    assert m["loss_abs"] <= 2e-6 and m["grad"] <= 1e-4, m
    m = G.mta_random_case(2, 160, [10, 5], 1, p=3.0)      # other channel count, general exponent
    assert m["loss_abs"] <= 2e-6 and m["grad"] <= 1e-4, m


def test_mta_edge_cases():
    import mm_distillnet_b200 as mmd
    from oracle import mmd_oracle as O
    crit = mmd.MTALoss("9", "2")
    # all-zero student features: the L2 norm is clamped at 1e-12 (F.normalize eps); B = 1; 1x1 level
    fs = [torch.zeros(1, 112, 4, 4), torch.randn(1, 112, 1, 1)]
    ft = [torch.randn(1, 112, 4, 4), torch.randn(1, 112, 1, 1)]
    ref = O.mta_loss([f.double() for f in fs], [f.double() for f in ft])
    out = crit([f.to(G.DEV) for f in fs], [f.to(G.DEV) for f in ft])
    assert torch.allclose(out.cpu().double(), ref, rtol=0, atol=2e-6)
    # helpers keep the reference signatures
    one = crit.mtaloss(fs[0].to(G.DEV) + 1.0, [ft[0].to(G.DEV), ft[0].to(G.DEV) * 2])
    ref1 = O.mta_level(fs[0].double() + 1.0, [ft[0].double(), ft[0].double() * 2])
    assert abs(one.item() - ref1.item()) < 2e-6
    at = crit.at(ft[0].to(G.DEV))
    assert H.rel_l2(at.cpu(), O.mta_at(ft[0])) < 1e-5
    with pytest.raises(RuntimeError):
        crit([f for f in fs], [f for f in ft])          # CPU tensors: no fallback
    with pytest.raises(TypeError):
        crit([f.to(G.DEV).half() for f in fs], [f.to(G.DEV).half() for f in ft])


def test_mta_c112_fast_path_matches_generic_kernels():
    """bf16 NHWC C=112 p=2 runs on the super-chunk kernels (mta_pool_c112 / mta_bwd_c112: 16 pixels = 7 coalesced warp
    loads); mmd_set_option("mta_fast", 0) switches back to the generic kernels: same losses, same gradients, including
    maps whose pixel count is no multiple of the 16-pixel super-chunk (18, 72, 2) and a single-pixel level."""
    import mm_distillnet_b200 as mmd
    from mm_distillnet_b200 import _lib
    gen = torch.Generator().manual_seed(3)
    sizes = [12, 6, 3, 1]
    mk = lambda: [(torch.randn(2, 112, s, s, generator=gen) * torch.exp(torch.randn(2, 1, s, s, generator=gen))).to(torch.bfloat16)
                  .to(G.DEV).contiguous(memory_format=torch.channels_last) for s in sizes]
    fs0, teachers = mk(), [mk() for _ in range(3)]
    go = torch.rand(3, len(sizes), generator=gen).to(G.DEV) * 0.01
    crit = mmd.MTALoss()

    def run():
        fs = [f.clone().requires_grad_(True) for f in fs0]
        each = crit.forward_each(fs, teachers)
        prod = crit(fs, teachers)
        ((each * go).sum() + 0.01 * prod.sum()).backward()
        return each.detach().clone(), prod.detach().clone(), [f.grad.clone() for f in fs]

    e1, p1, g1 = run()
    _lib.set_option("mta_fast", 0)
    try:
        e0, p0, g0 = run()
    finally:
        _lib.set_option("mta_fast", 1)
    assert torch.allclose(e1, e0, rtol=0, atol=1e-6) and torch.allclose(p1, p0, rtol=0, atol=1e-6)
    for a, b in zip(g1, g0):
        assert H.rel_l2(a.float().cpu(), b.float().cpu()) <= 1e-5       # same arithmetic up to fp32 summation order
