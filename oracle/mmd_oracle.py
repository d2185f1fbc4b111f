"""CPU oracle for the MM-DistillNet distillation hot path (TEST INFRASTRUCTURE ONLY).

This file is a plain-PyTorch (CPU, fp32 or fp64) restatement of the reference's algorithm for the
path named by BASELINE.json: the EfficientDet BiFPN cell / stack and the MTA loss.  It is the
checker the CUDA path is compared against; it is NOT part of the product.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may import it.
The product package (`mm_distillnet_b200`) never imports anything from `oracle/`.

Pinning: the reference ships no tests or golden vectors, so this restatement is
pinned against OUTPUTS OF THE REFERENCE ITSELF, run in the build container from /root/reference by
`oracle/make_golden.py`; the vectors live in `tests/golden/*.npz` and `tests/test_oracle_golden.py`
checks this file against them (forward values, input gradients, parameter gradients, BN running
statistics, MTA losses and MTA gradients).

Every function cites the reference lines it follows (paths relative to /root/reference).
Parameters are passed as a flat dict keyed by the reference's own state_dict names, so a reference
`state_dict()` can be used directly.
"""
import math

import torch
import torch.nn.functional as F

BN_MOMENTUM = 0.01   # src/YetAnotherEfficientDet.py:176, :239-265
BN_EPS = 1e-3        # same lines
FUSION_EPS = 1e-4    # src/YetAnotherEfficientDet.py:200 (epsilon=1e-4)

UP_NODES = ("conv6_up", "conv5_up", "conv4_up", "conv3_up")
DOWN_NODES = ("conv4_down", "conv5_down", "conv6_down", "conv7_down")
PROJ_NAMES = ("p5_down_channel", "p4_down_channel", "p3_down_channel", "p5_to_p6",
              "p4_down_channel_2", "p5_down_channel_2")


# ----------------------------------------------------------------------------------------------
# primitives
# ----------------------------------------------------------------------------------------------
def same_pad(size, kernel, stride):
    """TF-"SAME" padding split.  src/YetAnotherEfficientNet.py:54-60 (conv) and :93-99 (pool):
    extra = (ceil(size/stride)-1)*stride - size + kernel; before = extra//2; after = extra-before."""
    extra = (math.ceil(size / stride) - 1) * stride - size + kernel
    before = extra // 2
    return before, extra - before


def maxpool_same(x, hint=None):
    """MaxPool2dStaticSamePadding(3, 2): zero-pad (NOT -inf) then 3x3/s2 max.
    src/YetAnotherEfficientNet.py:90-104, instantiated at src/YetAnotherEfficientDet.py:228-231.

    `hint` (test-only, never part of the pinned path) forces the window arg-max to ANOTHER implementation's choice, so
    that both route the pooling gradients identically and the arg-max-flip discontinuity (a window whose two largest
    entries differ by less than the two implementations' rounding) drops out of gradient comparisons
    (tests/test_gpu_bifpn.py).  Forward values then differ from the true maximum by at most that top-2 gap.
      * floating tensor shaped like `x`: the other implementation's VALUES of the pooled tensor; the arg-max is taken on
        them (same zero padding, same first-maximum rule);
      * integer tensor [B, C, Ho, Wo]: the other implementation's recorded window indices, 0..8 row-major inside the
        3x3 window, 9 = the zero padding won (the encoding of the CUDA kernels' arg-max bytes)."""
    h, w = x.shape[-2:]
    top, bottom = same_pad(h, 3, 2)
    left, right = same_pad(w, 3, 2)
    x = F.pad(x, [left, right, top, bottom])
    if hint is None:
        return F.max_pool2d(x, 3, 2)
    wp = x.shape[-1]
    ho, wo = (x.shape[-2] - 3) // 2 + 1, (wp - 3) // 2 + 1
    if hint.is_floating_point():
        assert hint.shape[-2:] == (h, w) and hint.shape[:2] == x.shape[:2]
        hp = F.pad(hint.detach().to(torch.float64), [left, right, top, bottom])
        _, idx = F.max_pool2d(hp, 3, 2, return_indices=True)
        return x.flatten(-2).gather(-1, idx.flatten(-2)).view(idx.shape)
    assert tuple(hint.shape) == (x.shape[0], x.shape[1], ho, wo), (tuple(hint.shape), (x.shape[0], x.shape[1], ho, wo))
    k = hint.to(torch.int64)
    pad_won = k >= 9
    k = k.clamp(max=8)
    iy = 2 * torch.arange(ho).view(1, 1, ho, 1) + k // 3
    ix = 2 * torch.arange(wo).view(1, 1, 1, wo) + k % 3
    out = x.flatten(-2).gather(-1, (iy * wp + ix).flatten(-2)).view(k.shape)
    return torch.where(pad_won, torch.zeros((), dtype=x.dtype), out)


def upsample2(x):
    """nn.Upsample(scale_factor=2, mode='nearest'): out[y,x] = in[y//2, x//2].
    src/YetAnotherEfficientDet.py:223-226."""
    return x.repeat_interleave(2, dim=-2).repeat_interleave(2, dim=-1)


def swish(x):
    """MemoryEfficientSwish forward x*sigmoid(x); autograd of this expression equals the
    hand-written backward at src/YetAnotherEfficientNet.py:126-137."""
    return x * torch.sigmoid(x)


def batch_norm(x, p, prefix, training, stats_out=None):
    """nn.BatchNorm2d(momentum=0.01, eps=1e-3).  Train: batch mean / biased variance normalise,
    running <- 0.99*running + 0.01*(mean, unbiased var), num_batches_tracked += 1.
    Eval: running statistics.  src/YetAnotherEfficientDet.py:176,187 and :239-265.
    The running-stat update is returned through `stats_out` (dict) instead of mutating `p`."""
    w, b = p[prefix + ".weight"], p[prefix + ".bias"]
    rm, rv = p[prefix + ".running_mean"], p[prefix + ".running_var"]
    if training:
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
        if stats_out is not None:
            n = x.numel() // x.shape[1]
            with torch.no_grad():
                stats_out[prefix + ".running_mean"] = (1 - BN_MOMENTUM) * rm + BN_MOMENTUM * mean
                stats_out[prefix + ".running_var"] = (1 - BN_MOMENTUM) * rv + BN_MOMENTUM * var * (n / max(n - 1, 1))
                nbt = p.get(prefix + ".num_batches_tracked")
                if nbt is not None:
                    stats_out[prefix + ".num_batches_tracked"] = nbt + 1
    else:
        mean, var = rm, rv
    inv = torch.rsqrt(var + BN_EPS)
    return (x - mean[None, :, None, None]) * (inv * w)[None, :, None, None] + b[None, :, None, None]


def separable_block(x, p, prefix, training, stats_out=None, norm=True):
    """SeparableConvBlock.forward: depthwise 3x3 (no bias, SAME pad = 1,1,1,1) -> pointwise 1x1
    (+bias) -> BatchNorm (norm=True; no activation inside BiFPN; the detection heads build their blocks
    with norm=False and normalise per pyramid level themselves).  src/YetAnotherEfficientDet.py:154-192,
    padding src/YetAnotherEfficientNet.py:51-65."""
    c = x.shape[1]
    h, w = x.shape[-2:]
    top, bottom = same_pad(h, 3, 1)
    left, right = same_pad(w, 3, 1)
    x = F.pad(x, [left, right, top, bottom])
    x = F.conv2d(x, p[prefix + ".depthwise_conv.conv.weight"], None, groups=c)
    x = F.conv2d(x, p[prefix + ".pointwise_conv.conv.weight"], p[prefix + ".pointwise_conv.conv.bias"])
    if not norm:
        return x
    return batch_norm(x, p, prefix + ".bn", training, stats_out)


def projection(x, p, prefix, training, stats_out=None):
    """First-cell `pX_down_channel` = 1x1 conv (+bias) -> BatchNorm.
    src/YetAnotherEfficientDet.py:237-266."""
    x = F.conv2d(x, p[prefix + ".0.conv.weight"], p[prefix + ".0.conv.bias"])
    return batch_norm(x, p, prefix + ".1", training, stats_out)


def fusion_weights(w, eps=FUSION_EPS):
    """Fast normalised fusion: relu(w) / (sum(relu(w)) + eps).
    src/YetAnotherEfficientDet.py:338-339 (and the 7 sibling sites)."""
    w = F.relu(w)
    return w / (torch.sum(w, dim=0) + eps)


# ----------------------------------------------------------------------------------------------
# BiFPN cell / stack
# ----------------------------------------------------------------------------------------------
def bifpn_cell(inputs, p, prefix="", first_time=False, training=False, attention=True,
               eps=FUSION_EPS, stats_out=None, pool_hints=None):
    """One BiFPN cell.  src/YetAnotherEfficientDet.py:320-392 (attention=True) and :394-442
    (attention=False: plain sums).  `p` holds the cell's tensors under `prefix`.
    `pool_hints` (test-only, see maxpool_same): {"p3_out" | "p4_out" | "p5_out" | "p6_out" | (first cell) "p6_in" |
    "p7_in": tensor}, keyed by the name of the pooling's OUTPUT for the first-cell synthesis pools and of its INPUT for
    the bottom-up path."""
    q = prefix
    hint = (pool_hints or {}).get

    def fw(name, n):
        if attention:
            return fusion_weights(p[q + name], eps)
        return [1.0] * n

    if first_time:
        p3, p4, p5 = inputs
        p6_in = maxpool_same(projection(p5, p, q + "p5_to_p6", training, stats_out), hint("p6_in"))   # :324
        p7_in = maxpool_same(p6_in, hint("p7_in"))                                                    # :325
        p3_in = projection(p3, p, q + "p3_down_channel", training, stats_out)            # :327
        p4_in = projection(p4, p, q + "p4_down_channel", training, stats_out)            # :328
        p5_in = projection(p5, p, q + "p5_down_channel", training, stats_out)            # :329
    else:
        p3_in, p4_in, p5_in, p6_in, p7_in = inputs

    sep = lambda x, name: separable_block(x, p, q + name, training, stats_out)

    w = fw("p6_w1", 2)
    p6_up = sep(swish(w[0] * p6_in + w[1] * upsample2(p7_in)), "conv6_up")               # :338-341
    w = fw("p5_w1", 2)
    p5_up = sep(swish(w[0] * p5_in + w[1] * upsample2(p6_up)), "conv5_up")               # :344-347
    w = fw("p4_w1", 2)
    p4_up = sep(swish(w[0] * p4_in + w[1] * upsample2(p5_up)), "conv4_up")               # :350-353
    w = fw("p3_w1", 2)
    p3_out = sep(swish(w[0] * p3_in + w[1] * upsample2(p4_up)), "conv3_up")              # :356-359

    if first_time:
        p4_in = projection(p4, p, q + "p4_down_channel_2", training, stats_out)          # :362
        p5_in = projection(p5, p, q + "p5_down_channel_2", training, stats_out)          # :363

    w = fw("p4_w2", 3)
    p4_out = sep(swish(w[0] * p4_in + w[1] * p4_up + w[2] * maxpool_same(p3_out, hint("p3_out"))), "conv4_down")  # :366-370
    w = fw("p5_w2", 3)
    p5_out = sep(swish(w[0] * p5_in + w[1] * p5_up + w[2] * maxpool_same(p4_out, hint("p4_out"))), "conv5_down")  # :373-377
    w = fw("p6_w2", 3)
    p6_out = sep(swish(w[0] * p6_in + w[1] * p6_up + w[2] * maxpool_same(p5_out, hint("p5_out"))), "conv6_down")  # :380-384
    w = fw("p7_w2", 2)
    p7_out = sep(swish(w[0] * p7_in + w[1] * maxpool_same(p6_out, hint("p6_out"))), "conv7_down")        # :387-390
    return p3_out, p4_out, p5_out, p6_out, p7_out


def bifpn_stack(inputs, p, n_cells, prefix="", first_cell_first_time=True, training=False,
                attention=True, stats_out=None, pool_hints=None):
    """nn.Sequential(*[BiFPN(..., first_time=(i == 0), attention=...)]) as built at
    src/YetAnotherEfficientDet.py:639-644 and called at :668.  Keys are `<prefix><i>.<name>`.
    `pool_hints` (test-only): one dict per cell, see bifpn_cell."""
    feats = inputs
    for i in range(n_cells):
        feats = bifpn_cell(feats, p, prefix + "%d." % i, first_time=(i == 0 and first_cell_first_time),
                           training=training, attention=attention, stats_out=stats_out,
                           pool_hints=None if pool_hints is None else pool_hints[i])
    return feats


# ----------------------------------------------------------------------------------------------
# detection heads
# ----------------------------------------------------------------------------------------------
def head_tower(feat, p, prefix, level, num_layers, training, stats_out=None):
    """The shared-weight tower of Regressor / Classifier for one pyramid level: num_layers x
    [SeparableConvBlock(norm=False) -> the level's own BatchNorm -> swish].
    src/YetAnotherEfficientDet.py:466-470 (Regressor), :511-515 (Classifier)."""
    for i in range(num_layers):
        feat = separable_block(feat, p, prefix + "conv_list.%d" % i, training, norm=False)
        feat = batch_norm(feat, p, prefix + "bn_list.%d.%d" % (level, i), training, stats_out)
        feat = swish(feat)
    return feat


def regressor(inputs, p, prefix="", num_layers=3, training=False, stats_out=None):
    """Regressor.forward: per level tower -> header (dw3x3 -> pw 112 -> num_anchors*4) -> NHWC -> [B, H*W*A, 4];
    levels concatenated; second output = the last level's tower output.  src/YetAnotherEfficientDet.py:463-488."""
    feats, before_head = [], None
    for level, feat in enumerate(inputs):
        feat = head_tower(feat, p, prefix, level, num_layers, training, stats_out)
        before_head = feat
        feat = separable_block(feat, p, prefix + "header", training, norm=False)
        feat = feat.permute(0, 2, 3, 1).contiguous().view(feat.shape[0], -1, 4)
        feats.append(feat)
    return torch.cat(feats, dim=1), before_head


def classifier(inputs, p, num_anchors, num_classes, prefix="", num_layers=3, training=False, stats_out=None):
    """Classifier.forward: as the regressor with num_anchors*num_classes header channels viewed as
    [B, H*W*A, num_classes], concatenated over the levels, then sigmoid.  src/YetAnotherEfficientDet.py:508-532."""
    feats, before_head = [], None
    for level, feat in enumerate(inputs):
        feat = head_tower(feat, p, prefix, level, num_layers, training, stats_out)
        before_head = feat
        feat = separable_block(feat, p, prefix + "header", training, norm=False)
        feat = feat.permute(0, 2, 3, 1).contiguous()
        feat = feat.view(feat.shape[0], feat.shape[1], feat.shape[2], num_anchors, num_classes)
        feats.append(feat.contiguous().view(feat.shape[0], -1, num_classes))
    return torch.cat(feats, dim=1).sigmoid(), before_head


def synth_head_params(num_channels, out_channels, num_layers, seed, prefix="", n_levels=5, dtype=torch.float32):
    """Parameter / buffer dict of a Regressor (out_channels = num_anchors*4) or Classifier (num_anchors*num_classes)
    with the reference's state_dict names, filled from `synth`."""
    C = num_channels
    p = {}
    k = [seed * 1000 + 500]

    def nxt():
        k[0] += 1
        return k[0]

    for i in range(num_layers):
        pre = prefix + "conv_list.%d" % i
        p[pre + ".depthwise_conv.conv.weight"] = synth((C, 1, 3, 3), nxt(), 0.4, 0.0, dtype)
        p[pre + ".pointwise_conv.conv.weight"] = synth((C, C, 1, 1), nxt(), 1.5 / math.sqrt(C), 0.0, dtype)
        p[pre + ".pointwise_conv.conv.bias"] = synth((C,), nxt(), 0.1, 0.0, dtype)
    for lvl in range(n_levels):
        for i in range(num_layers):
            pre = prefix + "bn_list.%d.%d" % (lvl, i)
            p[pre + ".weight"] = synth((C,), nxt(), 0.5, 1.0, dtype)
            p[pre + ".bias"] = synth((C,), nxt(), 0.2, 0.0, dtype)
            p[pre + ".running_mean"] = synth((C,), nxt(), 0.3, 0.0, dtype)
            p[pre + ".running_var"] = synth((C,), nxt(), 0.5, 1.2, dtype)
            p[pre + ".num_batches_tracked"] = torch.tensor(3, dtype=torch.int64)
    pre = prefix + "header"
    p[pre + ".depthwise_conv.conv.weight"] = synth((C, 1, 3, 3), nxt(), 0.4, 0.0, dtype)
    p[pre + ".pointwise_conv.conv.weight"] = synth((out_channels, C, 1, 1), nxt(), 1.5 / math.sqrt(C), 0.0, dtype)
    p[pre + ".pointwise_conv.conv.bias"] = synth((out_channels,), nxt(), 0.5, -0.5, dtype)
    return p


# ----------------------------------------------------------------------------------------------
# MTA loss
# ----------------------------------------------------------------------------------------------
def mta_at(f, p=2.0):
    """MTALoss.at: F.normalize(f.pow(p).mean(1).view(B, -1)) (L2, dim=1, eps 1e-12).
    src/loss/MTALoss.py:76-77."""
    a = f.pow(p).mean(1).reshape(f.size(0), -1)
    return a / a.norm(p=2, dim=1, keepdim=True).clamp_min(1e-12)


def mta_level(f_s, f_t, T=9.0, p=2.0):
    """MTALoss.mtaloss for one pyramid level.  `f_t` is a tensor or a list of teacher tensors
    (product of attentions + L1 renormalisation when more than one).  The reference feeds
    PROBABILITIES (not log-probabilities) as kl_div's input; that quirk is kept:
    loss = sum_b sum_i t*(log t - s) / B.  src/loss/MTALoss.py:36-74."""
    a_s = mta_at(f_s, p)
    if torch.is_tensor(f_t):
        a_t = mta_at(f_t, p)
    elif len(f_t) == 1:
        a_t = mta_at(f_t[0], p)
    else:
        m = mta_at(f_t[0], p)
        for k in range(1, len(f_t)):
            m = m * mta_at(f_t[k], p)
        a_t = m / m.abs().sum(dim=1, keepdim=True).clamp_min(1e-12)
    s = torch.softmax(a_s / T, dim=1)
    t = torch.softmax(a_t / T, dim=1)
    return (torch.xlogy(t, t) - t * s).sum() / f_s.size(0)


def mta_loss(g_s, g_t, T=9.0, p=2.0):
    """MTALoss.forward: stack of per-level losses; `g_t` is a list of tensors (one teacher) or a
    list of per-teacher lists.  src/loss/MTALoss.py:15-34."""
    T, p = float(T), float(p)
    if torch.is_tensor(g_t[0]):
        return torch.stack([mta_level(fs, ft, T, p) for fs, ft in zip(g_s, g_t)], dim=0)
    return torch.stack([mta_level(g_s[i], [ft[i] for ft in g_t], T, p) for i in range(len(g_s))], dim=0)


def mta_grad_closed_form(f_s, a_t_normalised, grad_out, T=9.0):
    """Closed-form d(loss)/d(f_s) for p=2, used to cross-check autograd:
    g=-t/B; dz=s*(g-<g,s>); da_hat=dz/T; da=(da_hat - a_hat<a_hat,da_hat>)/||a||; df=(2/C) f da."""
    B, C = f_s.shape[:2]
    a = f_s.pow(2).mean(1).reshape(B, -1)
    nrm = a.norm(dim=1, keepdim=True).clamp_min(1e-12)
    ah = a / nrm
    s = torch.softmax(ah / T, dim=1)
    t = torch.softmax(a_t_normalised / T, dim=1)
    g = -t / B
    dz = s * (g - (g * s).sum(1, keepdim=True))
    dah = dz / T
    da = (dah - ah * (ah * dah).sum(1, keepdim=True)) / nrm
    return grad_out * (2.0 / C) * f_s * da.reshape(B, 1, *f_s.shape[2:])


# ----------------------------------------------------------------------------------------------
# detection loss
# ----------------------------------------------------------------------------------------------
def box_iou_anchor_gt(anchors, gt):
    """IoU of anchors [N,4] given as (y1, x1, y2, x2) against ground-truth boxes [M,4] given as (x1, y1, x2, y2):
    intersection / max(area_a + area_b - intersection, 1e-8).  src/loss/YetAnotherFocalLoss.py:6-20 (same operation order,
    so an fp32 run takes the same side of the 0.4 / 0.5 thresholds as the reference)."""
    area_b = (gt[:, 2] - gt[:, 0]) * (gt[:, 3] - gt[:, 1])
    iw = (torch.minimum(anchors[:, 3:4], gt[:, 2]) - torch.maximum(anchors[:, 1:2], gt[:, 0])).clamp(min=0)
    ih = (torch.minimum(anchors[:, 2:3], gt[:, 3]) - torch.maximum(anchors[:, 0:1], gt[:, 1])).clamp(min=0)
    inter = iw * ih
    union = (((anchors[:, 2] - anchors[:, 0]) * (anchors[:, 3] - anchors[:, 1])).unsqueeze(1) + area_b - inter).clamp(min=1e-8)
    return inter / union


def focal_loss(classifications, regressions, anchors, annotations, alpha=0.25, gamma=2.0, assign=None):
    """YetAnotherFocalLoss.forward (src/loss/YetAnotherFocalLoss.py:27-190) -> (regression_loss [1], classification_loss [1]).
    classifications [B,N,K] (sigmoid scores), regressions [B,N,4] (dy, dx, dh, dw), anchors [1,N,4] (y1,x1,y2,x2),
    annotations: list of B arrays [M_b,5] (x1, y1, x2, y2, class).  Per sample: anchors with max-IoU < 0.4 are negatives,
    >= 0.5 positives of the arg-max box's class, in between ignored; focal BCE summed and divided by max(#positives, 1);
    smooth-L1 (beta 1/9) on the positives' box deltas, mean over #positives*4; both averaged over the samples.  A sample
    without boxes contributes the all-negative classification sum (undivided) and a zero regression term; if NO sample has a
    box the reference skips every sample and returns zeros.
    `assign` (test-only): [B,N] int tensor from the other implementation (-2 ignore, -1 negative, m >= 0 positive of box m
    of the sample's valid boxes) used instead of this function's own IoU thresholds."""
    B = classifications.shape[0]
    dt = classifications.dtype
    a = anchors[0].to(dt)
    if max((len(x) for x in annotations), default=0) == 0:
        z = torch.zeros(1, dtype=dt)
        return z, z.clone()
    aw, ah = a[:, 3] - a[:, 1], a[:, 2] - a[:, 0]
    acx, acy = a[:, 1] + 0.5 * aw, a[:, 0] + 0.5 * ah
    cls_terms, reg_terms = [], []
    for b in range(B):
        c = classifications[b].clamp(1e-4, 1.0 - 1e-4)
        gt = torch.as_tensor(annotations[b], dtype=dt).reshape(-1, 5)
        gt = gt[gt[:, 4] != -1]
        neg_term = (1.0 - alpha) * c.pow(gamma) * (-torch.log(1.0 - c))
        if gt.shape[0] == 0:
            cls_terms.append(neg_term.sum())
            reg_terms.append(torch.zeros((), dtype=dt))
            continue
        if assign is None:
            iou_max, iou_arg = box_iou_anchor_gt(a, gt[:, :4]).max(dim=1)
            pos, neg = iou_max >= 0.5, iou_max < 0.4
        else:
            pos, neg, iou_arg = assign[b] >= 0, assign[b] == -1, assign[b].clamp(min=0).long()
        box = gt[iou_arg]
        onehot = torch.zeros_like(c, dtype=torch.bool)
        onehot[pos, box[pos, 4].long()] = True
        pos_term = alpha * (1.0 - c).pow(gamma) * (-torch.log(c))
        term = torch.where(onehot, pos_term, neg_term) * (pos | neg).unsqueeze(1).to(dt)
        npos = pos.sum()
        cls_terms.append(term.sum() / npos.to(dt).clamp(min=1.0))
        if int(npos) == 0:
            reg_terms.append(torch.zeros((), dtype=dt))
            continue
        box = box[pos]
        gw, gh = box[:, 2] - box[:, 0], box[:, 3] - box[:, 1]
        gcx, gcy = box[:, 0] + 0.5 * gw, box[:, 1] + 0.5 * gh
        gw, gh = gw.clamp(min=1), gh.clamp(min=1)
        tgt = torch.stack(((gcy - acy[pos]) / ah[pos], (gcx - acx[pos]) / aw[pos],
                           torch.log(gh / ah[pos]), torch.log(gw / aw[pos])), dim=1)
        d = (tgt - regressions[b][pos]).abs()
        reg_terms.append(torch.where(d <= 1.0 / 9.0, 0.5 * 9.0 * d * d, d - 0.5 / 9.0).mean())
    return torch.stack(reg_terms).mean(dim=0, keepdim=True), torch.stack(cls_terms).mean(dim=0, keepdim=True)


def synth_detections(B, N, K, seed, dtype=torch.float32):
    """Classification scores in (0, 1) with a few exact 0 / 1 / sub-clamp entries, and box deltas."""
    c = 0.5 + 0.5 * synth((B, N, K), seed, 1.0, 0.0, dtype)
    c = c.clamp(0.0, 1.0)
    flat = c.view(-1)
    flat[0], flat[1], flat[2], flat[3] = 0.0, 1.0, 5e-5, 1.0 - 5e-5
    r = synth((B, N, 4), seed + 1, 0.8, 0.0, dtype)
    return c, r


def synth_annotations(kind, B, size, K):
    """Deterministic pseudo-label sets for a `size` x `size` image: list of B float32 arrays [M_b, 5] (x1, y1, x2, y2, class).
    "mixed": 3 boxes / none / 1 box / 2 boxes ... (a sample without boxes in a batch that has some);
    "dense": 6 overlapping boxes per sample (arg-max competition between boxes); "none": no box in any sample."""
    import numpy as np
    out = []
    for b in range(B):
        if kind == "none" or (kind == "mixed" and b % 4 == 1):
            out.append(np.zeros((0, 5), dtype=np.float32))
            continue
        m = 6 if kind == "dense" else (3, 0, 1, 2)[b % 4]
        rows = []
        for j in range(m):
            t = 0.37 * (b + 1) + 0.91 * j
            cx, cy = size * (0.5 + 0.3 * math.sin(t)), size * (0.5 + 0.3 * math.cos(1.7 * t))
            w, h = size * (0.2 + 0.15 * math.sin(2.3 * t + 1.0) ** 2), size * (0.25 + 0.2 * math.cos(1.1 * t) ** 2)
            if kind == "dense":
                cx, cy = size * (0.45 + 0.04 * j), size * (0.5 + 0.03 * j)
            rows.append([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2, float((3 * b + 5 * j + 1) % K)])
        out.append(np.asarray(rows, dtype=np.float32))
    return out


# ----------------------------------------------------------------------------------------------
# pseudo-label generation
# ----------------------------------------------------------------------------------------------
def decode_boxes(anchors, regression):
    """YetAnotherEfficientDetBBoxTransform.forward (src/YetAnotherEfficientDet.py:574-602): anchors [1|B,N,4] (y1,x1,y2,x2),
    regression [B,N,4] (dy,dx,dh,dw) -> [B,N,4] (xmin, ymin, xmax, ymax)."""
    yc_a = (anchors[..., 0] + anchors[..., 2]) / 2
    xc_a = (anchors[..., 1] + anchors[..., 3]) / 2
    ha = anchors[..., 2] - anchors[..., 0]
    wa = anchors[..., 3] - anchors[..., 1]
    w = regression[..., 3].exp() * wa
    h = regression[..., 2].exp() * ha
    yc = regression[..., 0] * ha + yc_a
    xc = regression[..., 1] * wa + xc_a
    return torch.stack([xc - w / 2., yc - h / 2., xc + w / 2., yc + h / 2.], dim=2)


def nms_greedy(boxes, scores, iou_threshold):
    """torchvision.ops.nms (torchvision 0.26.0, the reference's dependency, not vendored in /root/reference), restated from
    its published CPU algorithm: candidates in STABLE descending score order; a kept box suppresses every later box whose
    IoU = inter / (area_i + area_j - inter) is > iou_threshold (areas (x2-x1)*(y2-y1), intersection sides clamped at 0).
    Returns the kept indices in score order.  boxes [n,4] (x1,y1,x2,y2)."""
    n = boxes.shape[0]
    if n == 0:
        return torch.zeros(0, dtype=torch.int64)
    order = torch.sort(scores, stable=True, descending=True).indices
    x1, y1, x2, y2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    areas = (x2 - x1) * (y2 - y1)
    dead = torch.zeros(n, dtype=torch.bool)
    keep = []
    for a in range(n):
        i = int(order[a])
        if dead[i]:
            continue
        keep.append(i)
        rest = order[a + 1:]
        w = (torch.minimum(x2[i], x2[rest]) - torch.maximum(x1[i], x1[rest])).clamp(min=0)
        h = (torch.minimum(y2[i], y2[rest]) - torch.maximum(y1[i], y1[rest])).clamp(min=0)
        inter = w * h
        dead[rest[inter / (areas[i] + areas[rest] - inter) > iou_threshold]] = True
    return torch.tensor(keep, dtype=torch.int64)


def batched_nms(boxes, scores, idxs, iou_threshold):
    """torchvision.ops.batched_nms, coordinate-trick branch (what it takes below 4 000 boxes on the CPU / 20 000 on CUDA):
    every class is moved to its own region by adding class * (max coordinate + 1) to its boxes, then one nms."""
    if boxes.numel() == 0:
        return torch.zeros(0, dtype=torch.int64)
    offsets = idxs.to(boxes) * (boxes.max() + torch.tensor(1).to(boxes))
    return nms_greedy(boxes + offsets[:, None], scores, iou_threshold)


def detections(classification, regression, anchors, valid_prediction_ids, conf_threshold, nms_threshold, image_size,
               ignore_labels=()):
    """EfficientDet_post_processing (src/utils/utils.py:144-231) up to the numeric result: per sample an array [n,6]
    (xmin, ymin, xmax, ymax, score, class) in NMS order.  Decode (:176), clip x/y minima at 0 and maxima at image_size
    (ClipBoxes :123-141), score = max class probability > conf_threshold (:178-179), arg-max class restricted to
    `valid_prediction_ids` (:197-204), class-wise NMS (:205), `ignore_labels` dropped (:212-215).
    Reference quirk kept: the reported score column is taken from the arg-max scores of ALL over-threshold anchors but
    indexed with positions in the class-FILTERED list (:195 vs :202-209), i.e. row j of the filtered list reports the score
    of the j-th over-threshold anchor; identical whenever no over-threshold anchor has a non-valid class."""
    boxes = decode_boxes(anchors[[0]], regression)
    boxes = torch.stack([boxes[..., 0].clamp(min=0), boxes[..., 1].clamp(min=0), boxes[..., 2].clamp(max=image_size),
                         boxes[..., 3].clamp(max=image_size)], dim=2)
    scores = classification.max(dim=2).values
    out = []
    valid = torch.tensor(sorted(valid_prediction_ids), dtype=torch.int64)
    for b in range(classification.shape[0]):
        over = scores[b] > conf_threshold
        if int(over.sum()) == 0:
            out.append(torch.zeros((0, 6), dtype=classification.dtype))
            continue
        all_scores, classes = classification[b, over].max(dim=1)
        bx, sc = boxes[b, over], scores[b, over]
        m = (classes[:, None] == valid[None, :]).any(-1)
        bx, cl, sc = bx[m], classes[m], sc[m]
        keep = batched_nms(bx, sc, cl, nms_threshold)
        if keep.numel() == 0:
            out.append(torch.zeros((0, 6), dtype=classification.dtype))
            continue
        cl_k, sc_k, bx_k = cl[keep], all_scores[keep], bx[keep]
        for lab in ignore_labels:
            sel = cl_k != lab
            bx_k, sc_k, cl_k = bx_k[sel], sc_k[sel], cl_k[sel]
        out.append(torch.cat([bx_k, sc_k[:, None], cl_k[:, None].to(bx_k)], dim=1))
    return out


def logits_to_ground_truth(logits, valid_prediction_ids, label_of_prediction, conf_threshold, nms_threshold, image_size,
                           ignore_labels=(), include_scores=False):
    """logits_to_ground_truth (src/utils/utils.py:234-324) with text_classes=False: per sample a float32 array of rows
    [int(max(xmin,0)), int(max(ymin,0)), int(min(xmax,S)), int(min(ymax,S)), (score,) label] where label =
    labels_txt2i[predictions_i2txt[class]] (here: `label_of_prediction[class]`)."""
    import numpy as np
    classification, regression, anchors = logits
    res = []
    for det in detections(classification, regression, anchors, valid_prediction_ids, conf_threshold, nms_threshold,
                          image_size, ignore_labels):
        rows = []
        for p in det.tolist():
            row = [int(max(p[0], 0)), int(max(p[1], 0)), int(min(p[2], image_size)), int(min(p[3], image_size))]
            if include_scores:
                row.append(p[4])
            row.append(label_of_prediction[int(p[5])])
            rows.append(row)
        res.append(np.array(rows, dtype=np.float32))
    return res


def merge_teacher_labels(per_teacher, iou_threshold=0.5, augment=False):
    """The cross-teacher integration of the step wrappers (src/optimization/train_methods.py:360-411):
    per sample, the teachers' [n,6] label arrays (with scores) are concatenated in teacher order, class-agnostic NMS at
    IoU 0.5 on the (integer-valued) boxes, the score column dropped, rows taken in NMS order.  A sample no teacher labelled
    stays an empty list.  `augment` (:384-386): when samples 0 and 1 both have rows, sample 1's list becomes sample 0's
    rows followed by its own before the NMS (sample 0 keeps its list)."""
    import numpy as np
    B = len(per_teacher[0])
    cats = []
    for b in range(B):
        rows = [np.asarray(t[b], dtype=np.float32).reshape(-1, 6) for t in per_teacher if np.asarray(t[b]).size > 0]
        cats.append(np.concatenate(rows, axis=0) if rows else None)
    if augment and B >= 2 and cats[0] is not None and cats[1] is not None:
        cats[1] = np.concatenate((cats[0], cats[1]), axis=0)
    out = []
    for b in range(B):
        cat = cats[b]
        if cat is None:
            out.append([])
            continue
        keep = nms_greedy(torch.from_numpy(cat[:, 0:4]), torch.from_numpy(cat[:, 4]), iou_threshold).numpy()
        out.append(np.delete(cat, 4, 1)[keep])
    return out


def synth_teacher_logits(B, anchors, K, seed, n_objects=6, size=128):
    """Detector-like outputs: every anchor gets a low background score; anchors overlapping one of a few pseudo objects get
    a high score for the object's class and box deltas that pull them towards it, so that thresholding + NMS has real work
    (clusters of overlapping over-threshold boxes, several classes, some over-threshold boxes of non-valid classes)."""
    N = anchors.shape[1]
    a = anchors[0]
    cls = 0.02 + 0.1 * (0.5 + 0.5 * synth((B, N, K), seed, 1.0, 0.0))
    reg = 0.05 * synth((B, N, 4), seed + 1, 1.0, 0.0)
    ha, wa = a[:, 2] - a[:, 0], a[:, 3] - a[:, 1]
    yc, xc = (a[:, 0] + a[:, 2]) / 2, (a[:, 1] + a[:, 3]) / 2
    for b in range(B):
        for j in range(n_objects if b % 3 != 2 else 0):          # every third sample: nothing above the threshold
            t = 0.77 * (b + 1) + 1.31 * j + 0.013 * seed
            ox, oy = size * (0.5 + 0.35 * math.sin(t)), size * (0.5 + 0.35 * math.cos(1.3 * t))
            ow, oh = size * (0.15 + 0.2 * math.sin(2.1 * t) ** 2), size * (0.15 + 0.25 * math.cos(0.7 * t) ** 2)
            k = (2 * b + 3 * j) % K
            gt = torch.tensor([[ox - ow / 2, oy - oh / 2, ox + ow / 2, oy + oh / 2]])
            iou = box_iou_anchor_gt(a, gt)[:, 0]
            hit = iou > 0.35
            cls[b, hit, k] = (0.25 + 0.7 * iou[hit]).clamp(max=0.97) + 0.01 * synth((int(hit.sum()),), seed + 7 * j + b, 1.0, 0.0)
            reg[b, hit, 0] = (oy - yc[hit]) / ha[hit] * 0.9
            reg[b, hit, 1] = (ox - xc[hit]) / wa[hit] * 0.9
            reg[b, hit, 2] = torch.log(oh / ha[hit]) * 0.9
            reg[b, hit, 3] = torch.log(ow / wa[hit]) * 0.9
    return cls.clamp(0.0, 1.0), reg


# ----------------------------------------------------------------------------------------------
# synthetic, RNG-free data so fixtures never depend on a torch/numpy RNG stream
# ----------------------------------------------------------------------------------------------
def synth(shape, seed, scale=1.0, offset=0.0, dtype=torch.float32):
    """Deterministic pseudo-random tensor from a closed form (no RNG): values in about [-1, 1]."""
    n = 1
    for s in shape:
        n *= s
    i = torch.arange(n, dtype=torch.float64)
    v = torch.sin(i * (0.618033988749895 + 0.0137 * seed) + 1.7 * seed) * 0.7 \
        + torch.sin(i * (2.39996322972865 + 0.0071 * seed) + 0.3 * seed) * 0.3
    return (v * scale + offset).reshape(shape).to(dtype)


def synth_cell_params(num_channels, conv_channels, first_time, seed, prefix="", dtype=torch.float32):
    """A full parameter/buffer dict for one cell with the reference's state_dict names and shapes
    (SURVEY.md A.4), filled from `synth` with values that exercise every term (non-trivial BN
    affine and running stats, fusion weights with a negative entry so the ReLU clamps)."""
    C = num_channels
    p = {}
    k = [seed * 1000]

    def nxt():
        k[0] += 1
        return k[0]

    def bn(pre):
        p[pre + ".weight"] = synth((C,), nxt(), 0.5, 1.0, dtype)
        p[pre + ".bias"] = synth((C,), nxt(), 0.2, 0.0, dtype)
        p[pre + ".running_mean"] = synth((C,), nxt(), 0.3, 0.0, dtype)
        p[pre + ".running_var"] = synth((C,), nxt(), 0.5, 1.2, dtype)
        p[pre + ".num_batches_tracked"] = torch.tensor(3, dtype=torch.int64)

    for name in UP_NODES + DOWN_NODES:
        pre = prefix + name
        p[pre + ".depthwise_conv.conv.weight"] = synth((C, 1, 3, 3), nxt(), 0.4, 0.0, dtype)
        p[pre + ".pointwise_conv.conv.weight"] = synth((C, C, 1, 1), nxt(), 1.5 / math.sqrt(C), 0.0, dtype)
        p[pre + ".pointwise_conv.conv.bias"] = synth((C,), nxt(), 0.1, 0.0, dtype)
        bn(pre + ".bn")
    if first_time:
        cin = {"p5_down_channel": conv_channels[2], "p4_down_channel": conv_channels[1],
               "p3_down_channel": conv_channels[0], "p5_to_p6": conv_channels[2],
               "p4_down_channel_2": conv_channels[1], "p5_down_channel_2": conv_channels[2]}
        for name in PROJ_NAMES:
            pre = prefix + name
            p[pre + ".0.conv.weight"] = synth((C, cin[name], 1, 1), nxt(), 1.5 / math.sqrt(cin[name]), 0.0, dtype)
            p[pre + ".0.conv.bias"] = synth((C,), nxt(), 0.1, 0.0, dtype)
            bn(pre + ".1")
    for name, n in (("p6_w1", 2), ("p5_w1", 2), ("p4_w1", 2), ("p3_w1", 2),
                    ("p4_w2", 3), ("p5_w2", 3), ("p6_w2", 3), ("p7_w2", 2)):
        p[prefix + name] = synth((n,), nxt(), 0.9, 0.8, dtype)   # in about [-0.1, 1.7]: some clamp at 0
    return p


def synth_stack_params(num_channels, conv_channels, n_cells, seed, prefix="", first_cell_first_time=True,
                       dtype=torch.float32):
    p = {}
    for i in range(n_cells):
        p.update(synth_cell_params(num_channels, conv_channels, i == 0 and first_cell_first_time,
                                   seed + i, prefix + "%d." % i, dtype))
    return p
