"""GPU parity of the detection loss (YetAnotherFocalLoss, src/loss/YetAnotherFocalLoss.py:23-190) through
the public module -> C ABI, against the stored outputs of the unmodified reference on its own anchors
(tests/golden/focal_*.npz) and against the fp64 CPU oracle at the D2 size (110 484 anchors).

Bounds: the anchor assignment (integer work) is bit-exact against the fp32 reference; losses <= 1e-5 relative, gradients
<= 1e-5 rel-L2 in fp32 storage (north_star: 1e-4); bf16 storage: the kernel computes in fp32 on the bf16-rounded inputs, so
against the oracle on the SAME rounded inputs the losses stay <= 1e-5 and the gradients within bf16 output rounding
(<= 4e-3 rel-L2)."""
import numpy as np
import pytest
import torch

import mm_distillnet_b200 as mmd
from oracle import mmd_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _run(c, r, anchors, ann, dtype, w_reg=1.3, w_cls=0.7):
    crit = mmd.YetAnotherFocalLoss()
    crit.record_assignment = True
    cd = c.to(DEV).to(dtype).requires_grad_(True)
    rd = r.to(DEV).to(dtype).requires_grad_(True)
    rl, cl = crit((cd, rd, anchors.to(DEV)), ann)
    out = {"reg_loss": rl.detach().cpu(), "cls_loss": cl.detach().cpu(), "assign": crit.last_assignment}
    if rl.requires_grad:
        (w_reg * rl + w_cls * cl).sum().backward()
        out["grad_cls"], out["grad_reg"] = cd.grad.float().cpu(), rd.grad.float().cpu()
    return out


def _oracle_assign(anchors, ann):
    """Anchor states of the fp32 oracle (== the fp32 reference): -2 ignored, -1 negative, m >= 0 positive of valid box m."""
    rows = []
    for a in ann:
        gt = torch.as_tensor(np.asarray(a, dtype=np.float32)).reshape(-1, 5)
        gt = gt[gt[:, 4] != -1]
        if gt.shape[0] == 0:
            rows.append(torch.full((anchors.shape[1],), -1, dtype=torch.int64))
            continue
        iou_max, iou_arg = O.box_iou_anchor_gt(anchors[0], gt[:, :4]).max(dim=1)
        st = torch.full_like(iou_arg, -2)
        st[iou_max < 0.4] = -1
        st[iou_max >= 0.5] = iou_arg[iou_max >= 0.5]
        rows.append(st)
    return torch.stack(rows)


@pytest.mark.parametrize("name", sorted(H.FOCAL_CASES))
def test_focal_golden_fp32(name):
    g = H.golden(name)
    c, r, anchors, ann = H.focal_case_inputs(name)
    out = _run(c, r, anchors, ann, torch.float32)
    assert out["reg_loss"].shape == (1,) and out["cls_loss"].shape == (1,)
    assert abs(float(out["reg_loss"]) - float(g["reg_loss"][0])) <= 1e-5 * max(1.0, abs(float(g["reg_loss"][0])))
    assert abs(float(out["cls_loss"]) - float(g["cls_loss"][0])) <= 1e-5 * max(1.0, abs(float(g["cls_loss"][0])))
    if "grad_cls" in g:
        assert torch.equal(out["assign"].cpu().long(), _oracle_assign(anchors, ann))       # bit-exact assignment
        assert H.rel_l2(out["grad_cls"], g["grad_cls"]) < 1e-5, H.rel_l2(out["grad_cls"], g["grad_cls"])
        assert H.rel_l2(out["grad_reg"], g["grad_reg"]) < 1e-5, H.rel_l2(out["grad_reg"], g["grad_reg"])
        assert np.array_equal(out["grad_reg"].numpy() == 0, g["grad_reg"] == 0)              # same set of positives
    else:
        assert "grad_cls" not in out and float(out["reg_loss"]) == 0.0 and float(out["cls_loss"]) == 0.0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_focal_full_size_vs_oracle(dtype):
    """The D2 problem size: 768x768 -> 110 484 anchors, 20 classes, B = 4 with 0 .. 12 boxes per sample, against the fp64
    oracle on the same (dtype-rounded) predictions with the CUDA assignment handed to the oracle — plus the assignment
    itself against the fp32 oracle, bit-exact."""
    size, K, B, A = 768, 20, 4, 9
    ys = []
    for lvl in range(3, 8):          # EfficientDet anchors (src/YetAnotherEfficientDet.py:71-152): stride 2^l, scale 4, 3 x 3 shapes
        stride = 2 ** lvl
        for sc in (2 ** 0, 2 ** (1.0 / 3.0), 2 ** (2.0 / 3.0)):
            for ra in ((1.0, 1.0), (1.4, 0.7), (0.7, 1.4)):
                half_x, half_y = 4.0 * stride * sc * ra[0] / 2.0, 4.0 * stride * sc * ra[1] / 2.0
                x = np.arange(stride / 2, size, stride)
                xv, yv = np.meshgrid(x, x)
                ys.append((lvl, np.stack((yv.reshape(-1) - half_y, xv.reshape(-1) - half_x, yv.reshape(-1) + half_y,
                                          xv.reshape(-1) + half_x), axis=1)))
    per_level = []
    for lvl in range(3, 8):
        boxes = [b for l, b in ys if l == lvl]
        per_level.append(np.stack(boxes, axis=1).reshape(-1, 4))
    anchors = torch.from_numpy(np.concatenate(per_level, axis=0).astype(np.float32)).unsqueeze(0)
    N = anchors.shape[1]
    assert N == 110484
    gen = torch.Generator().manual_seed(3)
    c = torch.rand(B, N, K, generator=gen).pow(3.0)                  # mostly small scores, like a trained classifier
    c.view(-1)[:4] = torch.tensor([0.0, 1.0, 5e-5, 1.0 - 5e-5])
    r = 0.5 * torch.randn(B, N, 4, generator=gen)
    ann = []
    for b, m in enumerate((12, 0, 1, 5)):
        xy = torch.rand(m, 2, generator=gen) * (size - 200)
        wh = 40 + torch.rand(m, 2, generator=gen) * 300
        cl = torch.randint(0, K, (m, 1), generator=gen).float()
        ann.append(torch.cat([xy, (xy + wh).clamp(max=size - 1), cl], dim=1).numpy().astype(np.float32))
    out = _run(c, r, anchors, ann, dtype)
    assign = out["assign"].cpu().long()
    assert torch.equal(assign, _oracle_assign(anchors, ann))
    assert int((assign >= 0).sum()) > 100 and int((assign == -2).sum()) > 100
    cq, rq = c.to(dtype).double().requires_grad_(True), r.to(dtype).double().requires_grad_(True)
    rl, cl_ = O.focal_loss(cq, rq, anchors.double(), ann, assign=assign)
    (1.3 * rl + 0.7 * cl_).sum().backward()
    assert abs(float(out["reg_loss"]) - float(rl)) <= 1e-5 * abs(float(rl))
    assert abs(float(out["cls_loss"]) - float(cl_)) <= 1e-5 * abs(float(cl_))
    tol = 1e-5 if dtype == torch.float32 else 4e-3
    assert H.rel_l2(out["grad_cls"], cq.grad) < tol, H.rel_l2(out["grad_cls"], cq.grad)
    assert H.rel_l2(out["grad_reg"], rq.grad) < tol, H.rel_l2(out["grad_reg"], rq.grad)


def test_focal_edge_cases():
    """No box anywhere -> zeros that do not depend on the predictions; padding rows and a sample without boxes inside a
    batch; one loss unused in the backward; CPU tensors and wrong shapes are refused."""
    c, r, anchors, ann = H.focal_case_inputs("focal_mixed")
    crit = mmd.YetAnotherFocalLoss()
    cd, rd = c.to(DEV).requires_grad_(True), r.to(DEV).requires_grad_(True)
    rl, cl = crit((cd, rd, anchors.to(DEV)), [np.zeros((0, 5), dtype=np.float32)] * c.shape[0])
    assert float(rl) == 0.0 and float(cl) == 0.0 and not rl.requires_grad
    # explicit -1 padding rows mean the same as shorter arrays
    padded = [np.concatenate([a.reshape(-1, 5), -np.ones((2, 5), dtype=np.float32)]) for a in ann]
    a1, a2 = _run(c, r, anchors, ann, torch.float32), _run(c, r, anchors, padded, torch.float32)
    assert torch.equal(a1["reg_loss"], a2["reg_loss"]) and torch.equal(a1["grad_reg"], a2["grad_reg"])
    assert abs(float(a1["cls_loss"]) - float(a2["cls_loss"])) <= 1e-6 * float(a1["cls_loss"])
    # only the classification loss is differentiated
    rl, cl = crit((cd, rd, anchors.to(DEV)), ann)
    cl.sum().backward()
    assert float(rd.grad.abs().max()) == 0.0 and float(cd.grad.abs().max()) > 0.0
    with pytest.raises(RuntimeError):
        crit((c, r, anchors), ann)
    with pytest.raises(ValueError):
        crit((cd, rd[:, :-1], anchors.to(DEV)), ann)
    with pytest.raises(ValueError):
        crit((cd, rd, anchors.to(DEV)), ann[:-1])
