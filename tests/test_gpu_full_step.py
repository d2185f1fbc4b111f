"""The whole distillation step behind the backbone, composed from the drop-ins and checked stage by stage against the oracle:

    ModelWithNMSLoss.forward (src/optimization/train_methods.py:425-516) around YetAnotherEfficientDet minus its backbone
    (src/YetAnotherEfficientDet.py:667-675: features = bifpn(p3, p4, p5); regression = regressor(features);
    classification = classifier(features)), the loss combination of src/optimization/traditional.py:171-182 and backward.

student (train): BiFPN stack -> Regressor / Classifier;  teachers (eval, no grad): the same three modules each;
teachers' predictions -> pseudo-labels (device) -> detection loss of the student; teachers' features -> MTA loss per teacher;
loss = w_main (reg + cls) + w_kd sum(kd);  backward into the backbone features and every student parameter.

The stages hand over at the points where a discrete decision is taken, so that a 1e-6 difference in a teacher score cannot
turn into another label set: teacher outputs (CUDA vs oracle, <= 2e-5); labels (device vs oracle ON the CUDA teacher outputs:
bit-exact); losses and gradients (CUDA vs fp64 oracle on those labels, with the CUDA anchor assignment and max-pool arg-max
handed to the oracle as in the loss / stack tests, <= 1e-4)."""
import numpy as np
import pytest
import torch
import torch.nn as nn

import mm_distillnet_b200 as mmd
from mm_distillnet_b200.bifpn import debug_pool_argmax
from oracle import mmd_oracle as O
from tests import helpers as H
from tests.test_gpu_heads import _build, rand_case

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
C, CC, A, K, L, N_CELLS = 112, [48, 120, 352], 9, 20, 3, 2
W_MAIN, W_KD = 1.0, 0.005


class _Det(nn.Module):
    """YetAnotherEfficientDet.forward behind the backbone (:667-675): (classification, regression, anchors), features."""

    def __init__(self, stack, reg, cls, anchors):
        super().__init__()
        self.bifpn, self.regressor, self.classifier, self.anchors = stack, reg, cls, anchors

    def forward(self, feats_in):
        features = self.bifpn(tuple(feats_in))
        regression, _ = self.regressor(features)
        classification, _ = self.classifier(features)
        return (classification, regression, self.anchors), features


class _Feats(list):
    """The backbone features of one network, passed where the wrappers expect an image batch (they read .device / .shape)."""

    @property
    def device(self):
        return self[0].device

    @property
    def shape(self):
        return self[0].shape


def _make_net(seed, train):
    torch.manual_seed(seed)
    stack = mmd.BiFPNStack(*[mmd.BiFPN(C, CC, first_time=(i == 0)) for i in range(N_CELLS)])
    sp = {k: v.clone() for k, v in stack.state_dict().items()}
    hp_r, _ = rand_case(C, A * 4, L, 1, 16, seed + 1)
    hp_c, _ = rand_case(C, A * K, L, 1, 16, seed + 2)
    hp_c["header.pointwise_conv.conv.bias"] = hp_c["header.pointwise_conv.conv.bias"] - 1.5      # sparse detections
    reg, cls = _build("reg", C, A, K, L, hp_r), _build("cls", C, A, K, L, hp_c)
    stack = stack.to(DEV)
    for m in (stack, reg, cls):
        m.train(train)
    return stack, reg, cls, sp, hp_r, hp_c


def _oracle_net(xs, sp, hp_r, hp_c, training, dt, hints=None, leaf_out=None):
    cast = (lambda v: v.to(dt) if v.is_floating_point() else v)
    if training:
        leaf = {k: (cast(v).requires_grad_(True) if v.is_floating_point() and "running" not in k else cast(v)) for k, v in sp.items()}
        lr = {k: (cast(v).requires_grad_(True) if v.is_floating_point() and "running" not in k else cast(v)) for k, v in hp_r.items()}
        lc = {k: (cast(v).requires_grad_(True) if v.is_floating_point() and "running" not in k else cast(v)) for k, v in hp_c.items()}
    else:
        leaf, lr, lc = ({k: cast(v) for k, v in d.items()} for d in (sp, hp_r, hp_c))
    if leaf_out is not None:
        leaf_out.update(bifpn=leaf, reg=lr, cls=lc)
    kw = {"pool_hints": hints} if hints is not None else {}
    f = O.bifpn_stack(tuple(xs), leaf, N_CELLS, first_cell_first_time=True, training=training, **kw)
    r, _ = O.regressor(tuple(f), lr, num_layers=L, training=training)
    c, _ = O.classifier(tuple(f), lc, A, K, num_layers=L, training=training)
    return c, r, f


def test_full_step_neck_heads_losses_fp32():
    B, s3, size, nt = 2, 32, 256, 2
    anchors = H.efficientdet_anchors(size)
    N = anchors.shape[1]
    assert N == 9 * sum((s3 >> i) ** 2 for i in range(5))
    gen = torch.Generator().manual_seed(21)

    def backbone_feats():
        return [torch.randn(B, c, s3 >> i, s3 >> i, generator=gen) for i, c in enumerate(CC)]
    x_s, x_t = backbone_feats(), [backbone_feats() for _ in range(nt)]
    s_stack, s_reg, s_cls, s_sp, s_hr, s_hc = _make_net(100, True)
    teachers = [_make_net(200 + 10 * t, False) for t in range(nt)]
    student = _Det(s_stack, s_reg, s_cls, anchors.to(DEV))
    tdict = nn.ModuleDict({m: _Det(t[0], t[1], t[2], anchors.to(DEV)) for m, t in zip(("rgb", "thermal"), teachers)})

    # ---- stage 1: teacher outputs, CUDA vs oracle (fp32 vs fp64) ----
    xt_d = [[x.to(DEV) for x in xs] for xs in x_t]
    t_out = []
    with torch.no_grad():
        for (m, det), xs in zip(tdict.items(), xt_d):
            (c, r, _), f = det(xs)
            t_out.append((c, r, f))
    for t in range(nt):
        co, ro, fo = _oracle_net([x.double() for x in x_t[t]], teachers[t][3], teachers[t][4], teachers[t][5], False, torch.float64)
        assert H.max_rel(t_out[t][0].cpu(), co) < 2e-5 and H.max_rel(t_out[t][1].cpu(), ro) < 2e-5
        for a, b in zip(t_out[t][2], fo):
            assert H.max_rel(a.float().cpu(), b) < 2e-5

    # ---- stage 2: labels, device vs oracle on the CUDA teacher outputs: bit-exact ----
    smax = torch.cat([c.max(dim=2).values.flatten() for c, _, _ in t_out])
    thr = float(torch.quantile(smax.float().cpu(), 1.0 - 0.015))           # ~1.5 % of the anchors fire
    cfg = H.pseudo_config(size, conf_threshold=repr(thr))
    vcd = H.pseudo_valid_classes_dict()
    label_of = {i: n for n, i in enumerate(H.PSEUDO_VALID_IDS)}
    crit_main, crit_kd = mmd.YetAnotherFocalLoss(), mmd.MTALoss("9", "2")
    crit_main.record_assignment = True
    model = mmd.ModelWithNMSLoss(student, tdict, crit_main, None, crit_kd, cfg, vcd)
    model.pseudo_max_rows, model.pseudo_max_labels = 1024, 2048
    x_sd = [x.to(DEV).requires_grad_(True) for x in x_s]
    pad = torch.zeros(B, 1, 2, 2, device=DEV)
    rgb, thermal = _Feats(xt_d[0]), _Feats(xt_d[1])
    out = model(rgb, thermal, pad, _Feats(x_sd), None)          # the fake "images" are the backbone features of each network
    labels = model.last_pseudo_labels.to_list()
    per_teacher = [O.logits_to_ground_truth((c.cpu(), r.cpu(), anchors), H.PSEUDO_VALID_IDS, label_of, conf_threshold=thr,
                                            nms_threshold=H.PSEUDO_CFG["nms_threshold"], image_size=size,
                                            ignore_labels=H.PSEUDO_CFG["ignore_labels"], include_scores=True) for c, r, _ in t_out]
    merged = O.merge_teacher_labels(per_teacher)
    n_rows = 0
    for b in range(B):
        ref = np.zeros((0, 5), dtype=np.float32) if len(merged[b]) == 0 else merged[b]
        got = np.zeros((0, 5), dtype=np.float32) if len(labels[b]) == 0 else labels[b]
        assert got.shape == ref.shape and np.array_equal(got, ref), b
        n_rows += ref.shape[0]
    assert n_rows >= 4

    # ---- stage 3: losses and gradients, CUDA vs fp64 oracle on those labels ----
    loss = W_MAIN * (out[0][0].mean() + out[1][0].mean()) + W_KD * torch.stack(out[2]).sum()      # traditional.py:171-182
    loss.backward()
    leafs = {}
    xo = [x.double().requires_grad_(True) for x in x_s]
    # the CUDA forward's own max-pool arg-max for the oracle's stack (tests/test_gpu_bifpn.py docstring); the forward is
    # deterministic, so a second pass over the same inputs records the same choices
    feats_again = student.bifpn(tuple(x.detach().requires_grad_(True) for x in x_sd))
    hints = debug_pool_argmax(feats_again[0])
    co, ro, fo = _oracle_net(xo, s_sp, s_hr, s_hc, True, torch.float64, hints=hints, leaf_out=leafs)
    rl, cl = O.focal_loss(co, ro, anchors.double(), merged, assign=crit_main.last_assignment.cpu().long())
    kd = torch.stack([O.mta_loss(list(fo), [f.double().cpu() for f in t_out[t][2]]) for t in range(nt)])
    (W_MAIN * (rl.mean() + cl.mean()) + W_KD * kd.sum()).backward()
    assert abs(float(out[0][0]) - float(rl)) <= 1e-4 * abs(float(rl)) and abs(float(out[1][0]) - float(cl)) <= 1e-4 * abs(float(cl))
    for t in range(nt):
        assert torch.allclose(out[2][t].detach().cpu().double(), kd[t], atol=2e-6, rtol=0)
    m = {}
    for i, (a, b) in enumerate(zip(x_sd, xo)):
        m["grad_in%d" % i] = H.rel_l2(a.grad.float().cpu(), b.grad)
    m["pgrad_bifpn"] = 0.0
    for k, p in student.bifpn.named_parameters():
        r = leafs["bifpn"][k].grad
        if k.endswith("conv.bias") or k[-3:-1] == "_w" or r is None or r.abs().max() == 0:
            continue
        m["pgrad_bifpn"] = max(m["pgrad_bifpn"], H.rel_l2(p.grad.float().cpu(), r))
    for name, mod in (("reg", student.regressor), ("cls", student.classifier)):
        worst = 0.0
        for k, p in mod.named_parameters():
            r = leafs[name][k].grad
            if r is None or r.abs().max() == 0 or (k.endswith("conv.bias") and not k.startswith("header")):
                continue           # a conv bias in front of a train-mode BatchNorm has a zero gradient up to rounding
            worst = max(worst, H.rel_l2(p.grad.float().cpu(), r))
        m["pgrad_" + name] = worst
    print("full step", {k: round(v, 7) for k, v in m.items()})
    bad = {k: v for k, v in m.items() if not v < 1e-4}
    assert not bad, (bad, m)
