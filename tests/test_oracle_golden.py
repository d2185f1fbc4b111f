"""Pin the oracle (oracle/mmd_oracle.py) against outputs of the real reference (tests/golden/)."""
import numpy as np
import pytest
import torch

from oracle import mmd_oracle as O
from tests import helpers as H

TOL = 2e-5   # oracle and reference are both fp32 torch-CPU; differences are summation-order only


@pytest.mark.parametrize("name", sorted(H.STACK_CASES))
def test_stack_matches_reference(name):
    C, cc, n_cells, first, B, s3, seed = H.STACK_CASES[name]
    g = H.golden(name)
    params, xs = H.stack_case_inputs(name)

    with torch.no_grad():
        ev = O.bifpn_stack(tuple(xs), params, n_cells, first_cell_first_time=first, training=False)
    for n, t in zip(H.LEVELS, ev):
        assert tuple(t.shape) == g["eval_" + n].shape
        assert H.max_rel(t, g["eval_" + n]) < TOL, (name, n)

    leaf = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v)
            for k, v in params.items()}
    xs = [x.requires_grad_(True) for x in xs]
    stats = {}
    tr = O.bifpn_stack(tuple(xs), leaf, n_cells, first_cell_first_time=first, training=True, stats_out=stats)
    gouts = H.stack_case_gouts(name, tr)
    sum((t * go).sum() for t, go in zip(tr, gouts)).backward()
    for n, t in zip(H.LEVELS, tr):
        # train-mode BN over a handful of values (P7 is 1x1, B=2) amplifies rounding: 1e-4 bar
        assert H.max_rel(t.detach(), g["train_" + n]) < 1e-4, (name, n)
    for i, x in enumerate(xs):
        assert H.rel_l2(x.grad, g["grad_in%d" % i]) < 2e-4, (name, i)
    for k, v in stats.items():
        ref = g["buf_" + k]
        if "num_batches" in k:
            assert int(v) == int(ref)
        else:
            assert H.max_rel(v, ref) < TOL, k
    for k, v in leaf.items():
        if not (torch.is_tensor(v) and v.requires_grad):
            continue
        if k.endswith("conv.bias"):
            # a bias feeding a train-mode BatchNorm has a mathematically zero gradient: rounding noise only
            continue
        if "pgrad_" + k in g:
            ref = g["pgrad_" + k]
            # fusion-weight gradients are differences of large sums: looser bar
            tol = 1e-2 if k[-3:-1] == "_w" else 5e-4
            assert H.rel_l2(v.grad, ref) < tol or np.abs(ref).max() < 1e-6, k
        else:
            s, nrm = g["pgsum_" + k]
            tol = 1e-2 if k[-3:-1] == "_w" else 5e-4
            assert abs(v.grad.double().norm().item() - nrm) <= tol * max(nrm, 1e-6), k


@pytest.mark.parametrize("name", sorted(H.PSEUDO_CASES))
def test_pseudo_labels_match_reference(name):
    """Pseudo-label restatement (decode, clip, threshold, class filter, class-wise NMS, label mapping, the
    cross-teacher NMS) against the outputs of the reference's own functions: identical rows in identical order."""
    B, size, K, seed, nt = H.PSEUDO_CASES[name]
    g = H.golden(name)
    anchors, logits, label_of = H.pseudo_case_inputs(name)
    per_teacher = []
    for t, (c, r) in enumerate(logits):
        got = O.logits_to_ground_truth((c, r, anchors), H.PSEUDO_VALID_IDS, label_of, image_size=size, include_scores=True,
                                       **H.PSEUDO_CFG)
        for b in range(B):
            ref = g["t%d_b%d" % (t, b)]
            assert got[b].shape == ref.shape and np.array_equal(got[b], ref), (t, b)
        per_teacher.append(got)
    assert sum(g["t0_b%d" % b].shape[0] for b in range(B)) >= 8            # the case has real detections ...
    assert any(g["t0_b%d" % b].size == 0 for b in range(B))                 # ... and a sample without any
    merged = O.merge_teacher_labels(per_teacher)
    for b in range(B):
        ref = g["merged_b%d" % b]
        got = np.zeros((0, 5), dtype=np.float32) if len(merged[b]) == 0 else merged[b]
        assert got.shape == ref.shape and np.array_equal(got, ref), b
    assert any(g["merged_b%d" % b].shape[0] < sum(g["t%d_b%d" % (t, b)].shape[0] for t in range(nt)) for b in range(B))
    # the augmented step's label merge (train_methods.py:384-386): sample 1 <- sample 0's rows + its own, before the NMS
    merged = O.merge_teacher_labels(per_teacher, augment=True)
    for b in range(B):
        ref = g["merged_aug_b%d" % b]
        got = np.zeros((0, 5), dtype=np.float32) if len(merged[b]) == 0 else merged[b]
        assert got.shape == ref.shape and np.array_equal(got, ref), ("augment", b)
    assert g["merged_aug_b1"].shape[0] > g["merged_b1"].shape[0] and np.array_equal(g["merged_aug_b0"], g["merged_b0"])


@pytest.mark.parametrize("name", sorted(H.FOCAL_CASES))
def test_focal_loss_matches_reference(name):
    """YetAnotherFocalLoss restatement against the reference's losses and gradients on its own anchors."""
    g = H.golden(name)
    c, r, anchors, ann = H.focal_case_inputs(name)
    c.requires_grad_(True)
    r.requires_grad_(True)
    rl, cl = O.focal_loss(c, r, anchors, ann)
    assert rl.shape == (1,) and cl.shape == (1,)
    assert abs(float(rl) - float(g["reg_loss"])) <= 1e-6 * max(1.0, abs(float(g["reg_loss"])))
    assert abs(float(cl) - float(g["cls_loss"])) <= 1e-6 * max(1.0, abs(float(g["cls_loss"])))
    if "grad_cls" in g:
        (1.3 * rl + 0.7 * cl).sum().backward()
        assert H.rel_l2(c.grad, g["grad_cls"]) < 1e-6 and H.rel_l2(r.grad, g["grad_reg"]) < 1e-6
    else:
        assert not rl.requires_grad and float(rl) == 0.0 and float(cl) == 0.0   # no box anywhere: the reference returns zeros


def test_focal_loss_forced_assignment_is_the_own_assignment():
    """The test-only `assign` argument (anchor states handed over by the CUDA implementation) reproduces the oracle's own
    thresholds when it is given the oracle's own states."""
    c, r, anchors, ann = H.focal_case_inputs("focal_dense")
    states = []
    for b in range(c.shape[0]):
        gt = torch.as_tensor(ann[b])
        iou_max, iou_arg = O.box_iou_anchor_gt(anchors[0], gt[:, :4]).max(dim=1)
        st = torch.full_like(iou_arg, -2)
        st[iou_max < 0.4] = -1
        st[iou_max >= 0.5] = iou_arg[iou_max >= 0.5]
        states.append(st)
    a = O.focal_loss(c, r, anchors, ann)
    b = O.focal_loss(c, r, anchors, ann, assign=torch.stack(states))
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert int((torch.stack(states) >= 0).sum()) > 20      # the case has positives of several boxes


@pytest.mark.parametrize("name", sorted(H.HEAD_CASES))
def test_heads_match_reference(name):
    """Regressor / Classifier restatement against outputs of the unmodified reference modules."""
    kind, C, A, K, L, B, s3, seed = H.HEAD_CASES[name]
    g = H.golden(name)
    params, xs = H.head_case_inputs(name)
    run = (lambda x, p, tr, st=None: O.regressor(x, p, num_layers=L, training=tr, stats_out=st)) if kind == "reg" else \
        (lambda x, p, tr, st=None: O.classifier(x, p, A, K, num_layers=L, training=tr, stats_out=st))
    with torch.no_grad():
        y, a = run(tuple(xs), params, False)
    assert tuple(y.shape) == g["eval_out"].shape and tuple(a.shape) == g["eval_align"].shape
    assert H.max_rel(y, g["eval_out"]) < TOL and H.max_rel(a, g["eval_align"]) < TOL
    leaf = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in params.items()}
    xs = [x.requires_grad_(True) for x in xs]
    stats = {}
    y, a = run(tuple(xs), leaf, True, stats)
    gy, ga = H.head_case_gouts(name, y, a)
    ((y * gy).sum() + (a * ga).sum()).backward()
    assert H.max_rel(y.detach(), g["train_out"]) < 1e-4 and H.max_rel(a.detach(), g["train_align"]) < 1e-4
    for i, x in enumerate(xs):
        assert H.rel_l2(x.grad, g["grad_in%d" % i]) < 2e-4, (name, i)
    for k, v in stats.items():
        if "num_batches" in k:
            assert int(v) == int(g["buf_" + k])
        else:
            assert H.max_rel(v, g["buf_" + k]) < TOL, k
    for k, v in leaf.items():
        if torch.is_tensor(v) and v.requires_grad:
            ref = g["pgrad_" + k]
            if k.startswith("conv_list") and k.endswith("conv.bias"):
                continue   # bias feeding a train-mode BatchNorm: zero gradient up to rounding
            assert H.rel_l2(v.grad, ref) < 5e-4 or np.abs(ref).max() < 1e-6, k


@pytest.mark.parametrize("name", sorted(H.MTA_CASES))
def test_mta_matches_reference(name):
    B, C, sizes, seed = H.MTA_CASES[name]
    g = H.golden(name)
    go = torch.tensor([0.005 * (i + 1) for i in range(len(sizes))])
    teachers = [H.structured_features(B, C, sizes, seed + 10 * (k + 1)) for k in range(3)]
    for branch, g_t in (("single", teachers[0]), ("multi", teachers)):
        g_s = [f.requires_grad_(True) for f in H.structured_features(B, C, sizes, seed)]
        loss = O.mta_loss(g_s, g_t, T="9", p="2")
        (loss * go).sum().backward()
        assert np.allclose(loss.detach().numpy(), g["loss_" + branch], rtol=0, atol=2e-6)
        for i, f in enumerate(g_s):
            assert H.rel_l2(f.grad, g["grad_%s_%d" % (branch, i)]) < 1e-4, (branch, i)
    # fp64 oracle agrees with the fp64 reference run to rounding
    l64 = O.mta_loss([f.double() for f in H.structured_features(B, C, sizes, seed)],
                     [f.double() for f in teachers[0]])
    assert np.allclose(l64.numpy(), g["loss_single_fp64"], rtol=0, atol=1e-12)


def test_mta_closed_form_gradient():
    """SURVEY.md A.3 closed form (what the CUDA backward implements) == autograd, in fp64."""
    B, C, sizes, seed = H.MTA_CASES["mta_c16"]
    fs = [f.double().requires_grad_(True) for f in H.structured_features(B, C, sizes, seed)]
    ft = [f.double() for f in H.structured_features(B, C, sizes, seed + 10)]
    loss = O.mta_loss(fs, ft)
    loss.sum().backward()
    for f, t in zip(fs, ft):
        cf = O.mta_grad_closed_form(f.detach(), O.mta_at(t), 1.0)
        assert H.rel_l2(cf, f.grad) < 1e-10


def test_landmarks():
    """SURVEY.md 8c numeric landmarks: near-uniform attention -> loss = -ln(HW) - 1/HW."""
    torch.manual_seed(0)
    for s in (6, 12, 24):
        fs, ft = torch.randn(2, 112, s, s), torch.randn(2, 112, s, s)
        l = O.mta_level(fs, ft).item()
        n = s * s
        assert abs(l - (-np.log(n) - 1.0 / n)) < 2e-3
    w = O.fusion_weights(torch.ones(2))
    assert abs(w[0].item() - 0.499975) < 1e-6
    w = O.fusion_weights(torch.ones(3))
    assert abs(w[0].item() - 0.333322) < 1e-6
    x = -torch.ones(1, 1, 4, 4)
    assert O.maxpool_same(x).flatten().tolist() == [-1.0, 0.0, 0.0, 0.0]   # zero pad wins at the border


def test_pool_hint_with_own_values_is_the_pinned_path():
    """maxpool_same(x, hint=x) (the test-only arg-max forcing used by the fp32 GPU gradient checks) equals the pinned
    maxpool_same(x) in value and gradient, including ties against the zero padding and first-maximum ties."""
    x = O.synth((2, 5, 9, 12), 3).requires_grad_(True)
    with torch.no_grad():
        x[0, 0, :2, :2] = 0.25          # a 4-way tie inside one window: the first maximum takes the gradient
        x[1, :, -1, :] = -1.0           # bottom windows: the padding zero wins
    g = O.synth((2, 5, 5, 6), 4)
    a = O.maxpool_same(x)
    (ga,) = torch.autograd.grad((a * g).sum(), x)
    b = O.maxpool_same(x, hint=x.detach().clone())
    (gb,) = torch.autograd.grad((b * g).sum(), x)
    assert torch.equal(a, b) and torch.equal(ga, gb)
    # a hint that prefers another element of a window reroutes the gradient and moves the value by the top-2 gap only
    h = x.detach().clone()
    h[0, 1, 4, 4] += 10.0
    c = O.maxpool_same(x, hint=h)
    (gc,) = torch.autograd.grad((c * g).sum(), x)
    assert gc[0, 1, 4, 4] != ga[0, 1, 4, 4] and (c - a).abs().max() <= 2.0


def test_pool_hint_as_window_indices():
    """Integer hints (the CUDA kernels' arg-max bytes: 0..8 inside the window, 9 = padding) select exactly those elements."""
    x = O.synth((1, 3, 6, 6), 5).requires_grad_(True)
    with torch.no_grad():
        x[0, 0, 4:, :] = -1.0       # the bottom windows of channel 0 hold only negative values: the padding zero wins
    ref = O.maxpool_same(x)
    # derive the indices of the natural arg-max, then check the index path reproduces value and gradient
    xp = torch.nn.functional.pad(x.detach(), [0, 1, 0, 1])
    _, flat = torch.nn.functional.max_pool2d(xp, 3, 2, return_indices=True)
    iy, ix = flat // 7, flat % 7
    oy = 2 * torch.arange(3).view(1, 1, 3, 1)
    ox = 2 * torch.arange(3).view(1, 1, 1, 3)
    k = (iy - oy) * 3 + (ix - ox)
    k = torch.where((iy >= 6) | (ix >= 6), torch.full_like(k, 9), k).to(torch.uint8)
    got = O.maxpool_same(x, hint=k)
    assert torch.equal(got, ref)
    g = O.synth((1, 3, 3, 3), 6)
    (ga,) = torch.autograd.grad((ref * g).sum(), x, retain_graph=True)
    (gb,) = torch.autograd.grad((got * g).sum(), x)
    assert torch.equal(ga, gb)
    assert (k == 9).any()          # the bottom-row windows of channel 0 are won by the padding


def test_stack_pool_hints_reproduce_unhinted_stack():
    name = "stack2_c16_odd"
    C, cc, n_cells, first, B, s3, seed = H.STACK_CASES[name]
    params, xs = H.stack_case_inputs(name)
    with torch.no_grad():
        feats, hints = tuple(xs), []
        for i in range(n_cells):   # per-cell outputs of the same implementation as hints
            feats = O.bifpn_cell(feats, params, "%d." % i, first_time=(i == 0 and first), training=True)
            hints.append(dict(zip(("p3_out", "p4_out", "p5_out", "p6_out"), feats)))
        a = O.bifpn_stack(tuple(xs), params, n_cells, first_cell_first_time=first, training=True)
        b = O.bifpn_stack(tuple(xs), params, n_cells, first_cell_first_time=first, training=True, pool_hints=hints)
    for u, v in zip(a, b):
        assert torch.equal(u, v)
