"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules from /root/reference.

Run in the build container only (the GPU box has no /root/reference):
    python oracle/make_golden.py
Inputs and weights are produced by the RNG-free closed forms in oracle/mmd_oracle.py (`synth*`), so
the fixtures store only the reference's OUTPUTS (forward values, gradients, BN buffers).  Tests
regenerate the inputs with the same closed forms and compare the oracle (CPU tests) and the CUDA
path (GPU tests) against these stored reference outputs.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import types  # noqa: E402
sys.modules.setdefault("cv2", types.ModuleType("cv2"))   # the reference's loss file imports cv2 without using it; not installed here

from oracle import mmd_oracle as O  # noqa: E402
from src.YetAnotherEfficientDet import BiFPN, Classifier, Regressor  # noqa: E402  (reference)
from src.loss.MTALoss import MTALoss  # noqa: E402  (reference)
from src.loss.YetAnotherFocalLoss import YetAnotherFocalLoss  # noqa: E402  (reference)
from src.YetAnotherEfficientDet import Anchors  # noqa: E402  (reference)

OUT = os.path.join(ROOT, "tests", "golden")
# name -> (annotation kind, B, image size, classes, seed)   (mirrored by tests/helpers.py FOCAL_CASES)
FOCAL_CASES = {"focal_mixed": ("mixed", 4, 128, 20, 21), "focal_dense": ("dense", 2, 128, 20, 22), "focal_none": ("none", 2, 128, 20, 23)}
LEVEL_NAMES = ("p3", "p4", "p5", "p6", "p7")


def load_cell(cell, params, prefix):
    sd = {k[len(prefix):]: v.clone() for k, v in params.items() if k.startswith(prefix)}
    missing, unexpected = cell.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    return cell


def pyramid_inputs(B, C, s3, seed, dtype=torch.float32):
    sizes = [s3, s3 // 2, s3 // 4, s3 // 8, max(s3 // 16, 1)]
    return [O.synth((B, C, s, s), seed + i, 1.0, 0.1 * i, dtype) for i, s in enumerate(sizes)]


def backbone_inputs(B, conv_channels, s3, seed, dtype=torch.float32):
    return [O.synth((B, c, s3 >> i, s3 >> i), seed + i, 1.0, 0.0, dtype) for i, c in enumerate(conv_channels)]


def run_stack_case(name, C, conv_channels, n_cells, first, B, s3, seed, save_param_grads):
    """Reference nn.Sequential of BiFPN cells: eval forward, train forward + backward."""
    params = O.synth_stack_params(C, conv_channels, n_cells, seed, first_cell_first_time=first)
    cells = [load_cell(BiFPN(C, conv_channels, first_time=(i == 0 and first)), params, "%d." % i)
             for i in range(n_cells)]
    stack = torch.nn.Sequential(*cells)
    out = {}
    mk = (lambda: backbone_inputs(B, conv_channels, s3, seed + 50)) if first else \
        (lambda: pyramid_inputs(B, C, s3, seed + 50))

    stack.eval()
    with torch.no_grad():
        ev = stack(tuple(mk()))
    for n, t in zip(LEVEL_NAMES, ev):
        out["eval_" + n] = t.numpy()

    stack.train()
    xs = [x.requires_grad_(True) for x in mk()]
    tr = stack(tuple(xs))
    gouts = [O.synth(tuple(t.shape), seed + 70 + i, 1.0, 0.0) for i, t in enumerate(tr)]
    loss = sum((t * g).sum() for t, g in zip(tr, gouts))
    loss.backward()
    for n, t in zip(LEVEL_NAMES, tr):
        out["train_" + n] = t.detach().numpy()
    for i, x in enumerate(xs):
        out["grad_in%d" % i] = x.grad.numpy()
    sd = stack.state_dict()
    for k, v in sd.items():
        if "running_" in k or "num_batches" in k:
            out["buf_" + k] = v.numpy()
    for k, v in stack.named_parameters():
        if save_param_grads:   # full tensors: the tests compare element-wise (a permuted / sign-flipped tensor fails)
            out["pgrad_" + k] = v.grad.numpy()
        if save_param_grads != "only":   # (sum, norm) per parameter, kept for the older norm-only checks
            out["pgsum_" + k] = np.array([v.grad.double().sum().item(), v.grad.double().norm().item()])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "->", len(out), "arrays")


def run_head_case(name, kind, C, num_anchors, num_classes, num_layers, B, s3, seed):
    """Reference Regressor / Classifier: eval forward, train forward + backward (both outputs carry a gradient)."""
    out_ch = num_anchors * (4 if kind == "reg" else num_classes)
    params = O.synth_head_params(C, out_ch, num_layers, seed)
    head = Regressor(C, num_anchors, num_layers) if kind == "reg" else Classifier(C, num_anchors, num_classes, num_layers)
    missing, unexpected = head.load_state_dict({k: v.clone() for k, v in params.items()}, strict=True)
    assert not missing and not unexpected
    out = {}
    head.eval()
    with torch.no_grad():
        y, a = head(tuple(pyramid_inputs(B, C, s3, seed + 50)))
    out["eval_out"], out["eval_align"] = y.numpy(), a.numpy()
    head.train()
    xs = [x.requires_grad_(True) for x in pyramid_inputs(B, C, s3, seed + 50)]
    y, a = head(tuple(xs))
    gy = O.synth(tuple(y.shape), seed + 70, 1.0, 0.0)
    ga = O.synth(tuple(a.shape), seed + 71, 1.0, 0.0)
    ((y * gy).sum() + (a * ga).sum()).backward()
    out["train_out"], out["train_align"] = y.detach().numpy(), a.detach().numpy()
    for i, x in enumerate(xs):
        out["grad_in%d" % i] = x.grad.numpy()
    for k, v in head.state_dict().items():
        if "running_" in k or "num_batches" in k:
            out["buf_" + k] = v.numpy()
    for k, v in head.named_parameters():
        out["pgrad_" + k] = v.grad.numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "->", len(out), "arrays")


def structured_features(B, C, sizes, seed):
    """Features whose channel-pooled attention is far from uniform (SURVEY.md 8d 'structured' set)."""
    fs = []
    for i, s in enumerate(sizes):
        base = O.synth((B, C, s, s), seed + i, 1.0, 0.0)
        mod = torch.exp(1.5 * O.synth((B, 1, s, s), seed + 100 + i, 1.0, 0.0))
        fs.append(base * mod)
    return fs


def run_mta_case(name, B, C, sizes, seed):
    out = {}
    crit = MTALoss(T="9", p="2")   # strings, as extract_criterions_from_config passes them
    g_s = [f.requires_grad_(True) for f in structured_features(B, C, sizes, seed)]
    teachers = [structured_features(B, C, sizes, seed + 10 * (k + 1)) for k in range(3)]
    go = torch.tensor([0.005 * (i + 1) for i in range(len(sizes))])

    loss1 = crit(g_s, teachers[0])                      # shipped-cfg branch (MTALoss.py:17-19)
    (loss1 * go).sum().backward()
    out["loss_single"] = loss1.detach().numpy()
    for i, f in enumerate(g_s):
        out["grad_single_%d" % i] = f.grad.numpy().copy()
        f.grad = None

    loss3 = crit(g_s, teachers)                         # multi-teacher product branch (:20-34)
    (loss3 * go).sum().backward()
    out["loss_multi"] = loss3.detach().numpy()
    for i, f in enumerate(g_s):
        out["grad_multi_%d" % i] = f.grad.numpy().copy()
        f.grad = None

    crit_d = MTALoss()                                  # fp64 landmark for the loss values
    l64 = crit_d([f.detach().double() for f in g_s], [f.double() for f in teachers[0]])
    out["loss_single_fp64"] = l64.numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "->", len(out), "arrays")


PSEUDO_VALID_IDS = [2, 4, 6, 7, 9, 11, 14, 16, 19]    # prediction ids of the classes kept (mirrored by tests/helpers.py)
PSEUDO_CASES = {"pseudo_a": (4, 128, 20, 100, 3)}       # name -> (B, image size, classes, seed, teachers)


def reference_pseudo_label_functions():
    """EfficientDet_post_processing / logits_to_ground_truth / ClipBoxes exactly as written in the reference, compiled from
    their lines of src/utils/utils.py (:123-324) — the module itself does not import here (librosa, tensorboardX,
    hpbandster, ... are not installed) and none of its other 2 300 lines is on this path."""
    import torch.nn as nn
    from torchvision.ops import nms
    from torchvision.ops.boxes import batched_nms
    from src.YetAnotherEfficientDet import YetAnotherEfficientDetBBoxTransform
    txt = open("/root/reference/src/utils/utils.py").read()
    sl = txt[txt.index("class ClipBoxes"):txt.index("def filter_model_dict")]
    ns = {"torch": torch, "nn": nn, "np": np, "batched_nms": batched_nms, "nms": nms,
          "YetAnotherEfficientDetBBoxTransform": YetAnotherEfficientDetBBoxTransform, "EfficientDetBBoxTransform": None}
    exec(compile(sl, "/root/reference/src/utils/utils.py[123:324]", "exec"), ns)
    return ns, nms


def run_pseudo_case(name, B, size, K, seed, n_teachers):
    """Reference pseudo-label generation: logits_to_ground_truth(include_scores=True) per teacher (train_methods.py:343-349)
    and the cross-teacher integration (:360-411, restated here line by line around torchvision's nms, as it is inline
    code of a forward method)."""
    import configparser
    ns, nms = reference_pseudo_label_functions()
    cfg = configparser.ConfigParser()
    cfg.read_dict({"s": {"conf_threshold": "0.3", "nms_threshold": "0.5", "image_size": str(size),
                         "student": "YetAnotherEfficientDet", "ignore_labels": "4"}})
    vcd = {"predictions_txt2i": {"c%d" % i: i for i in PSEUDO_VALID_IDS}, "predictions_i2txt": {i: "c%d" % i for i in PSEUDO_VALID_IDS},
           "labels_txt2i": {"c%d" % i: n for n, i in enumerate(PSEUDO_VALID_IDS)}}
    anchors = Anchors(anchor_scale=4.)(torch.zeros(1, 3, size, size), torch.float32)
    out = {"anchors": anchors.numpy()}
    batch_labels = [[] for _ in range(B)]
    for t in range(n_teachers):
        c, r = O.synth_teacher_logits(B, anchors, K, seed + 10 * t, size=size)
        this = ns["logits_to_ground_truth"](logits=(c, r, anchors), anchors=None, valid_classes_dict=vcd, config=cfg["s"],
                                            include_scores=True)
        for b in range(B):
            out["t%d_b%d" % (t, b)] = np.asarray(this[b], dtype=np.float32)
            if np.asarray(this[b]).size == 0:
                continue
            batch_labels[b] = this[b] if len(batch_labels[b]) == 0 else np.concatenate((batch_labels[b], this[b]), axis=0)
    for tag, augment in (("merged", False), ("merged_aug", True)):
        labels = [x if len(x) == 0 else x.copy() for x in batch_labels]
        if augment and len(labels[1]) != 0 and len(labels[0]) != 0:          # train_methods.py:384-386
            labels[1] = np.concatenate((labels[0], labels[1]), axis=0)
        for b in range(B):
            if len(labels[b]) == 0:
                out["%s_b%d" % (tag, b)] = np.zeros((0, 5), dtype=np.float32)
                continue
            idx = nms(boxes=torch.from_numpy(labels[b][:, 0:4]), scores=torch.from_numpy(labels[b][:, 4]),
                      iou_threshold=0.5).numpy()
            out["%s_b%d" % (tag, b)] = np.delete(labels[b], 4, 1)[idx]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "per teacher:", [[out["t%d_b%d" % (t, b)].shape[0] for b in range(B)] for t in range(n_teachers)],
          "merged:", [out["merged_b%d" % b].shape[0] for b in range(B)])


def run_focal_case(name, kind, B, size, K, seed):
    """Reference YetAnotherFocalLoss on the reference's own Anchors for a size x size image: losses and the gradients of
    1.3 * regression_loss + 0.7 * classification_loss w.r.t. the classification scores and box deltas."""
    anchors = Anchors(anchor_scale=4.)(torch.zeros(1, 3, size, size), torch.float32)
    N = anchors.shape[1]
    c, r = O.synth_detections(B, N, K, seed)
    ann = O.synth_annotations(kind, B, size, K)
    c.requires_grad_(True)
    r.requires_grad_(True)
    rl, cl = YetAnotherFocalLoss()((c, r, anchors), ann)
    out = {"anchors": anchors.numpy(), "reg_loss": rl.detach().numpy(), "cls_loss": cl.detach().numpy()}
    if rl.requires_grad or cl.requires_grad:
        (1.3 * rl + 0.7 * cl).sum().backward()
        out["grad_cls"], out["grad_reg"] = c.grad.numpy(), r.grad.numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "N=%d" % N, "reg %.6f cls %.4f" % (float(rl), float(cl)), "annots", [len(a) for a in ann])


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(4)
    # small-channel cases pin every term of the oracle, including all parameter gradients
    run_stack_case("cell_c16", 16, [8, 12, 20], 1, False, B=2, s3=16, seed=1, save_param_grads="only")
    run_stack_case("first_c16", 16, [8, 12, 20], 1, True, B=2, s3=16, seed=2, save_param_grads="only")
    run_stack_case("stack3_c16", 16, [8, 12, 20], 3, True, B=2, s3=32, seed=3, save_param_grads="only")
    # odd top level (P7 = 3x3) : P3=48 .. P6=6
    run_stack_case("stack2_c16_odd", 16, [8, 12, 20], 2, True, B=1, s3=48, seed=4, save_param_grads="only")
    # D2 channel counts (what the CUDA kernels are built for), tiny spatial size; full parameter gradients as well
    run_stack_case("stack2_c112", 112, [48, 120, 352], 2, True, B=2, s3=16, seed=5, save_param_grads=True)
    run_stack_case("cell_c112", 112, [48, 120, 352], 1, False, B=2, s3=16, seed=6, save_param_grads=True)
    # pseudo-label generation
    for name, (B, size, K, seed, nt) in PSEUDO_CASES.items():
        run_pseudo_case(name, B, size, K, seed, nt)
    # detection loss: the reference's own anchors for a 128x128 image (3 069 boxes), 20 classes
    for name, (kind, B, size, K, seed) in FOCAL_CASES.items():
        run_focal_case(name, kind, B, size, K, seed)
    # detection heads: D2 configuration (112 channels, 9 anchors, 3 layers; 20 classes), tiny pyramid
    run_head_case("reg_c112", "reg", 112, 9, 20, 3, B=2, s3=16, seed=9)
    run_head_case("cls_c112", "cls", 112, 9, 20, 3, B=2, s3=16, seed=10)
    run_mta_case("mta_c112", B=2, C=112, sizes=[12, 6, 3], seed=7)
    run_mta_case("mta_c16", B=3, C=16, sizes=[16, 8, 4, 2, 1], seed=8)


if __name__ == "__main__":
    main()
