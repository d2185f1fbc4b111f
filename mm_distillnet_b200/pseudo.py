"""Pseudo-label generation on device: drop-ins for the reference's `logits_to_ground_truth`
(src/utils/utils.py:234-324, with EfficientDet_post_processing :144-231 inside) and for the cross-teacher integration of
the step wrappers (src/optimization/train_methods.py:186-250, :343-411).

The reference turns every teacher's `(classification, regression, anchors)` into Python lists per sample (`.cpu()` per
sample and teacher, torchvision NMS on small tensors, numpy concatenation, one more NMS) and copies the result back for
the detection loss.  Here `teacher_pseudo_labels()` is ONE call of `mmd_pseudo_labels` (4 launches for all teachers and
samples) that leaves a padded `[B, M, 5]` annotation tensor on the device — `YetAnotherFocalLoss` takes it as is, so a
training step never synchronises the host.  `logits_to_ground_truth()` keeps the reference's signature and return type
(list of numpy arrays) for callers that want the lists (one device→host copy at the end).

CUDA only; there is no CPU fallback.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

DEFAULT_CAP = 4096          # over-threshold anchors per (teacher, sample)
DEFAULT_MAX_ROWS = 256      # rows per (teacher, sample) after the class-wise NMS
DEFAULT_MAX_LABELS = 256    # merged rows per sample
OVERFLOW_BITS = {1: "over-threshold anchors beyond `cap`", 2: "rows per teacher beyond `max_rows`",
                 4: "merged rows beyond `max_labels`"}


class PseudoLabels(object):
    """Device-resident result.  `boxes` float32 [B, M, 5] = (x1, y1, x2, y2, label) in NMS order, rows >= counts[b] are -1
    (the reference's annot_padded, YetAnotherFocalLoss.py:35-39); `counts` int32 [B + 1] ([B] = overflow bit mask);
    `teacher_rows` float32 [T, B, R, 6] / `teacher_counts` int32 [T, B] = the per-teacher rows (with scores)."""

    def __init__(self, boxes, counts, teacher_rows, teacher_counts):
        self.boxes, self.counts, self.teacher_rows, self.teacher_counts = boxes, counts, teacher_rows, teacher_counts

    def __len__(self):
        return self.boxes.shape[0]

    def check_overflow(self):
        """Host synchronisation: raises if a capacity was exceeded (rows were dropped in score order)."""
        bits = int(self.counts[-1].item())
        if bits:
            raise RuntimeError("pseudo-label capacities exceeded: " + "; ".join(v for k, v in OVERFLOW_BITS.items() if bits & k))

    def to_list(self):
        """The reference's `batch_labels`: per sample a float32 array [n, 5], or [] for a sample nobody labelled
        (train_methods.py:314, :395-397).  Synchronises."""
        self.check_overflow()
        counts = self.counts[:-1].cpu().numpy()
        boxes = self.boxes.cpu().numpy()
        return [boxes[b, :int(n)].copy() if n > 0 else [] for b, n in enumerate(counts)]

    def teacher_lists(self):
        """Per teacher the return value of the reference's logits_to_ground_truth(include_scores=True): a list of B float32
        arrays [n, 6] (shape (0,) when empty, as `np.array([], dtype=np.float32)` is)."""
        self.check_overflow()
        counts = self.teacher_counts.cpu().numpy()
        rows = self.teacher_rows.cpu().numpy()
        return [[rows[t, b, :int(counts[t, b])].copy() if counts[t, b] > 0 else np.array([], dtype=np.float32)
                 for b in range(rows.shape[1])] for t in range(rows.shape[0])]


def _cfg_get(config, key, kind, default=None):
    """configparser section (the reference's `config`, getfloat / getint) or a plain dict."""
    if hasattr(config, "getfloat") and not isinstance(config, dict):
        if key not in config:
            if default is None:
                raise KeyError(key)
            return default
        return {"float": config.getfloat, "int": config.getint}.get(kind, config.get)(key)
    if key not in config:
        if default is None:
            raise KeyError(key)
        return default
    v = config[key]
    return float(v) if kind == "float" else int(v) if kind == "int" else v


def label_table(valid_classes_dict, num_classes):
    """int32 [K]: label id of prediction id k = labels_txt2i[predictions_i2txt[k]] (utils.py:299-300) for the prediction ids
    in predictions_txt2i.values() (:197-201), -1 for every other id."""
    tab = np.full(num_classes, -1, dtype=np.int32)
    for k in valid_classes_dict["predictions_txt2i"].values():
        k = int(k)
        if 0 <= k < num_classes:
            tab[k] = int(valid_classes_dict["labels_txt2i"][valid_classes_dict["predictions_i2txt"][k]])
    if (tab[tab >= 0] < 0).any():
        raise ValueError("label ids must be >= 0")
    return tab


_TABLE_CACHE = {}
_WS_CACHE = {}


def _device_table(tab, dev):
    key = (tab.tobytes(), str(dev))
    if key not in _TABLE_CACHE:
        _TABLE_CACHE[key] = torch.from_numpy(tab.copy()).to(dev)
    return _TABLE_CACHE[key]


def teacher_pseudo_labels(predictions, valid_classes_dict, config, cap=DEFAULT_CAP, max_rows=DEFAULT_MAX_ROWS,
                          max_labels=DEFAULT_MAX_LABELS, raw_rows=False, merge_iou=0.5, merge_batch_0_1=False):
    """predictions: list (one per teacher, in teacher order) of `(classification [B,N,K], regression [B,N,4], anchors
    [1|B,N,4])` — the first element of what the reference's models return.  Runs, for every teacher,
    logits_to_ground_truth(include_scores=True) and then the wrappers' integration; returns PseudoLabels (device).
    CUDA-graph capture: call once eagerly first (the class-id table and the workspace are created on the first call; the
    captured call then only allocates its outputs from the graph's pool).
    `merge_batch_0_1=True` is the augmented step's label merge (train_methods.py:384-386): when samples 0 and 1 both have
    labels, sample 1 gets sample 0's rows in front of its own before the cross-teacher NMS."""
    if len(predictions) < 1 or len(predictions) > _lib.PL_MAX_TEACHERS:
        raise ValueError("1..%d teachers, got %d" % (_lib.PL_MAX_TEACHERS, len(predictions)))
    cls0, reg0, anchors = predictions[0][0], predictions[0][1], predictions[0][2]
    if not cls0.is_cuda:
        raise RuntimeError("mm_distillnet_b200 pseudo-label generation needs CUDA tensors (there is no CPU fallback)")
    if cls0.dim() != 3 or cls0.dtype not in (torch.float32, torch.bfloat16):
        raise TypeError("classification must be float32 or bfloat16 [B,N,K], got %s %s" % (cls0.dtype, tuple(cls0.shape)))
    B, N, K = cls0.shape
    dev = cls0.device
    a = _lib.PseudoArgs()
    a.B, a.N, a.K, a.T = B, N, K, len(predictions)
    a.dtype = _lib.MMD_F32 if cls0.dtype == torch.float32 else _lib.MMD_BF16
    a.cap, a.max_rows, a.max_labels, a.raw_rows = int(cap), int(max_rows), int(max_labels), 1 if raw_rows else 0
    a.conf_threshold = _cfg_get(config, "conf_threshold", "float")
    a.nms_threshold = _cfg_get(config, "nms_threshold", "float")
    a.image_size = float(_cfg_get(config, "image_size", "int"))
    a.merge_iou = float(merge_iou)
    a.merge01 = 1 if merge_batch_0_1 else 0
    ignore = _cfg_get(config, "ignore_labels", "str", default="")
    ignore = [int(x) for x in ignore.split(",") if x.strip()] if isinstance(ignore, str) else [int(x) for x in ignore]
    if len(ignore) > _lib.PL_MAX_IGNORE:
        raise ValueError("at most %d ignore_labels" % _lib.PL_MAX_IGNORE)
    a.n_ignore = len(ignore)
    for i, v in enumerate(ignore):
        a.ignore[i] = v
    keep = []
    for t, pred in enumerate(predictions):
        c, r = pred[0], pred[1]
        if c.shape != (B, N, K) or r.shape != (B, N, 4) or c.dtype != cls0.dtype or r.dtype != cls0.dtype or c.device != dev:
            raise ValueError("teacher %d: expected classification %s / regression %s of %s" % (t, (B, N, K), (B, N, 4), cls0.dtype))
        c, r = c.detach().contiguous(), r.detach().contiguous()
        keep += [c, r]
        a.cls[t], a.reg[t] = c.data_ptr(), r.data_ptr()
    if anchors.dim() != 3 or anchors.shape[1] != N or anchors.shape[2] != 4:
        raise ValueError("anchors must be [1,N,4] with N=%d, got %s" % (N, tuple(anchors.shape)))
    anc = anchors[0].detach().to(device=dev, dtype=torch.float32).contiguous()      # `anchors[[0]]`, utils.py:168
    tab = _device_table(label_table(valid_classes_dict, K), dev)
    a.anchors, a.label_of = anc.data_ptr(), tab.data_ptr()
    need = int(_lib.lib().mmd_pseudo_workspace_bytes(C.byref(a)))
    wkey = (str(dev), torch.cuda.current_stream(dev).cuda_stream)
    ws = _WS_CACHE.get(wkey)
    if ws is None or ws.numel() < need:
        ws = _WS_CACHE[wkey] = torch.empty(need, dtype=torch.uint8, device=dev)
    rows = torch.empty((a.T, B, a.max_rows, 6), dtype=torch.float32, device=dev)
    tcounts = torch.empty((a.T, B), dtype=torch.int32, device=dev)
    labels = torch.empty((B, a.max_labels, 5), dtype=torch.float32, device=dev)
    counts = torch.empty(B + 1, dtype=torch.int32, device=dev)
    a.workspace, a.teacher_rows, a.teacher_counts = ws.data_ptr(), rows.data_ptr(), tcounts.data_ptr()
    a.labels, a.counts = labels.data_ptr(), counts.data_ptr()
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().mmd_pseudo_labels(C.byref(a), torch.cuda.current_stream().cuda_stream), "mmd_pseudo_labels")
    del keep
    return PseudoLabels(labels, counts, rows, tcounts)


def logits_to_ground_truth(logits, anchors, valid_classes_dict, config, include_scores=False, crash_if_no_pred=False,
                           text_classes=False, regressBoxes=None, clipBoxes=None):
    """Signature and return value of src/utils/utils.py:234-324 (`anchors` is unused there as well: the anchors come with
    `logits`).  `text_classes=True` returns lists of [xmin, ymin, xmax, ymax, (score,) class name] like the reference;
    custom `regressBoxes` / `clipBoxes` modules are not supported (the YetAnotherEfficientDet transform and ClipBoxes at
    image_size are what the kernels restate)."""
    if regressBoxes is not None or clipBoxes is not None:
        raise NotImplementedError("custom regressBoxes / clipBoxes: only YetAnotherEfficientDetBBoxTransform + ClipBoxes(image_size)")
    if "student" in config and "YetAnotherEfficientDet" not in config["student"]:
        raise NotImplementedError("only the YetAnotherEfficientDet box encoding is built (config['student'] = %r)" % config["student"])
    out = teacher_pseudo_labels([logits], valid_classes_dict, config, raw_rows=text_classes)
    res = []
    image_size = _cfg_get(config, "image_size", "int")
    for rows in out.teacher_lists()[0]:
        if crash_if_no_pred:
            assert len(rows) > 0
        if text_classes:
            preds = []
            for p in rows.reshape(-1, 6).tolist():
                box = [int(max(p[0], 0)), int(max(p[1], 0)), int(min(p[2], image_size)), int(min(p[3], image_size))]
                name = valid_classes_dict["predictions_i2txt"][int(p[5])]
                preds.append(box + ([p[4]] if include_scores else []) + [name])
            res.append(preds)
        elif len(rows) == 0:
            res.append(np.array([], dtype=np.float32))
        else:
            res.append(rows if include_scores else np.delete(rows, 4, 1))
    return res
