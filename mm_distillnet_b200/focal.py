"""Drop-in for the reference's detection loss (src/loss/YetAnotherFocalLoss.py:23-190) on one sm_100a kernel per direction.

`YetAnotherFocalLoss()(prediction, annotations)` keeps the reference's call: `prediction = (classifications [B,N,K],
regressions [B,N,4], anchors [1,N,4])`, `annotations` = a list of B numpy arrays `[M_b, 5]` (x1, y1, x2, y2, class) — or a
device-resident `PseudoLabels` (pseudo.py: the teachers' labels, never copied to the host) — and returns `(regression_loss [1], classification_loss [1])`.  The whole batch is ONE launch (`mmd_focal_fwd`): the labels
are padded on the host and copied once, the IoU matrix, the assignment and the per-sample temporaries never exist in
HBM.  CUDA only; predictions float32 or bfloat16, loss values float32.  No CPU fallback.
"""
import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .pseudo import PseudoLabels

ALPHA, GAMMA = 0.25, 2.0     # src/loss/YetAnotherFocalLoss.py:44-45


def pad_annotations(annotations):
    """list of [M_b, 5] arrays -> float32 [B, max M_b, 5] padded with -1 (the reference's annot_padded, :35-39)."""
    m = max((int(np.shape(a)[0]) for a in annotations), default=0)
    out = np.full((len(annotations), m, 5), -1.0, dtype=np.float32)
    for i, a in enumerate(annotations):
        if np.shape(a)[0] > 0:
            out[i, :np.shape(a)[0], :] = np.asarray(a, dtype=np.float32).reshape(-1, 5)
    return out


class _FocalFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cls, reg, anchors, boxes, want_assign):
        B, N, K = cls.shape
        dev = cls.device
        a = _lib.FocalArgs()
        a.B, a.N, a.K, a.M = B, N, K, boxes.shape[1]
        a.dtype = _lib.MMD_F32 if cls.dtype == torch.float32 else _lib.MMD_BF16
        a.alpha, a.gamma = ALPHA, GAMMA
        cls_c, reg_c = cls.detach().contiguous(), reg.detach().contiguous()
        acc = torch.zeros(B * 4, dtype=torch.float64, device=dev)
        loss = torch.empty(2, dtype=torch.float32, device=dev)
        assign = torch.empty((B, N), dtype=torch.int32, device=dev) if want_assign else None
        a.cls, a.reg, a.anchors, a.boxes = cls_c.data_ptr(), reg_c.data_ptr(), anchors.data_ptr(), boxes.data_ptr()
        a.acc, a.loss = acc.data_ptr(), loss.data_ptr()
        a.assign = assign.data_ptr() if assign is not None else None
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().mmd_focal_fwd(C.byref(a), stream), "mmd_focal_fwd")
        ctx.args, ctx.keep, ctx.dev = a, (cls_c, reg_c, anchors, boxes, acc, loss), dev
        if assign is None:
            assign = torch.empty(0, dtype=torch.int32, device=dev)
        ctx.mark_non_differentiable(assign)
        return loss[0:1], loss[1:2], assign

    @staticmethod
    def backward(ctx, g_reg, g_cls, _g_assign):
        cls_c, reg_c = ctx.keep[0], ctx.keep[1]
        gr = None if g_reg is None else g_reg.detach().to(torch.float32).contiguous()
        gc = None if g_cls is None else g_cls.detach().to(torch.float32).contiguous()
        grad_cls, grad_reg = torch.empty_like(cls_c), torch.empty_like(reg_c)
        with torch.cuda.device(ctx.dev):
            stream = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().mmd_focal_bwd(C.byref(ctx.args), None if gr is None else gr.data_ptr(),
                                                None if gc is None else gc.data_ptr(), grad_cls.data_ptr(), grad_reg.data_ptr(),
                                                stream), "mmd_focal_bwd")
        return grad_cls, grad_reg, None, None, None


class YetAnotherFocalLoss(nn.Module):
    """Signature and semantics of src/loss/YetAnotherFocalLoss.py:23-190 (alpha 0.25, gamma 2, IoU thresholds 0.4 / 0.5,
    smooth-L1 beta 1/9, EfficientDet box encoding)."""

    def __init__(self):
        super(YetAnotherFocalLoss, self).__init__()
        self.last_assignment = None     # TEST HOOK: [B, N] anchor states of the last call when `record_assignment` is set
        self.record_assignment = False

    def forward(self, prediction, annotations, **kwargs):
        classifications, regressions, anchors = prediction
        if not classifications.is_cuda:
            raise RuntimeError("mm_distillnet_b200.YetAnotherFocalLoss needs CUDA tensors (there is no CPU fallback)")
        if classifications.dtype not in (torch.float32, torch.bfloat16) or regressions.dtype != classifications.dtype:
            raise TypeError("classifications / regressions must both be float32 or bfloat16")
        if classifications.dim() != 3 or regressions.shape != classifications.shape[:2] + (4,):
            raise ValueError("expected classifications [B,N,K] and regressions [B,N,4], got %s / %s"
                             % (tuple(classifications.shape), tuple(regressions.shape)))
        B, N, K = classifications.shape
        if len(annotations) != B:
            raise ValueError("%d annotation arrays for a batch of %d" % (len(annotations), B))
        if anchors.dim() != 3 or anchors.shape[1] != N or anchors.shape[2] != 4:
            raise ValueError("anchors must be [1,N,4] with N=%d, got %s" % (N, tuple(anchors.shape)))
        dev = classifications.device
        if isinstance(annotations, PseudoLabels):
            # labels made on this device by mmd_pseudo_labels: already the padded [B, M, 5] tensor; the "no box in any
            # sample" case (zeros, :61-62 / :181-188) is decided inside the kernels, so nothing synchronises the host
            boxes = annotations.boxes
            if boxes.device != dev or boxes.dtype != torch.float32 or not boxes.is_contiguous():
                raise ValueError("PseudoLabels.boxes must be a contiguous float32 tensor on %s" % dev)
            if boxes.shape[1] > _lib.FOCAL_MAX_BOXES:
                raise ValueError("at most %d rows per sample, got %d" % (_lib.FOCAL_MAX_BOXES, boxes.shape[1]))
        else:
            padded = pad_annotations(annotations)
            if padded.shape[1] == 0:
                # no box in any sample: the reference skips every sample (:61-62) and returns zeros that do not depend on
                # the predictions (:181-188)
                z = torch.zeros(1, dtype=torch.float32, device=dev)
                return z, z.clone()
            if padded.shape[1] > _lib.FOCAL_MAX_BOXES:
                raise ValueError("at most %d boxes per sample, got %d" % (_lib.FOCAL_MAX_BOXES, padded.shape[1]))
            boxes = torch.from_numpy(padded).to(dev, non_blocking=True)
        anc = anchors[0].detach().to(device=dev, dtype=torch.float32).contiguous()
        reg_loss, cls_loss, assign = _FocalFunction.apply(classifications, regressions, anc, boxes, self.record_assignment)
        self.last_assignment = assign if self.record_assignment else None
        return reg_loss, cls_loss
