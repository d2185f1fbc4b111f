"""Drop-in for the reference's `MTALoss` (src/loss/MTALoss.py:9-77) on hand-written sm_100a kernels.

Same constructor (`MTALoss(T=9.0, p=2.0)`, strings accepted as `extract_criterions_from_config` passes them,
src/utils/utils.py:1603-1604), same `forward(g_s, g_t) -> Tensor[len(g_s)]`, same `.mtaloss` / `.at` helpers.
All levels and all teachers of one call go through ONE `mmd_mta_fwd` (three launches) and the backward through
ONE `mmd_mta_bwd` launch.  CUDA only: there is no CPU fallback.
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib


def _layout_of(tensors):
    """Pick the layout the kernels read without a copy; otherwise convert everything to channels_last."""
    if all(t.is_contiguous(memory_format=torch.channels_last) for t in tensors):
        return _lib.MMD_NHWC, tensors
    if all(t.is_contiguous() for t in tensors):
        return _lib.MMD_NCHW, tensors
    return _lib.MMD_NHWC, [t.contiguous(memory_format=torch.channels_last) for t in tensors]


def _check(feats):
    f0 = feats[0]
    if not f0.is_cuda:
        raise RuntimeError("mm_distillnet_b200.MTALoss needs CUDA tensors (there is no CPU fallback)")
    if f0.dtype not in (torch.float32, torch.bfloat16):
        raise TypeError("MTALoss supports float32 and bfloat16 features (fp16 underflows the MTA gradients), got %s" % f0.dtype)
    for t in feats:
        if t.dim() != 4:
            raise ValueError("MTALoss expects [B,C,H,W] feature maps, got shape %s" % (tuple(t.shape),))
        if t.device != f0.device:
            raise RuntimeError("MTALoss: features live on different devices")


class _MTAFunction(torch.autograd.Function):
    """loss[l] for all levels; feats = n_levels student maps followed by n_teachers * n_levels teacher maps.

    `separate` (bool): n_teachers independent single-teacher calls sharing the student maps -> loss[n_teachers, n_levels]
    (MmdMtaArgs.separate); otherwise one call against the product of the teachers -> loss[n_levels]."""

    @staticmethod
    def forward(ctx, T, p, n_levels, n_teachers, separate, *feats):
        _check(feats)
        dtype = feats[0].dtype
        feats = [f.detach() if f.dtype == dtype else f.detach().to(dtype) for f in feats]
        layout, feats = _layout_of(feats)
        fs = feats[:n_levels]
        B, Cch = fs[0].shape[0], fs[0].shape[1]
        if Cch % 4 != 0:
            raise ValueError("MTALoss kernels need a channel count that is a multiple of 4, got %d" % Cch)
        for l in range(n_levels):
            for k in range(n_teachers + 1):
                t = feats[k * n_levels + l]
                if tuple(t.shape) != tuple(fs[l].shape):
                    raise ValueError("MTALoss: level %d shapes differ: %s vs %s" % (l, tuple(t.shape), tuple(fs[l].shape)))
        dev = fs[0].device
        sum_hw = sum(f.shape[2] * f.shape[3] for f in fs)
        need_grad = any(ctx.needs_input_grad[5:5 + n_levels])
        ncalls = n_teachers if separate else 1
        att = torch.empty((1 + n_teachers) * B * sum_hw, dtype=torch.float32, device=dev)
        ga = torch.empty(ncalls * B * sum_hw, dtype=torch.float32, device=dev) if need_grad else None
        loss_b = torch.empty(ncalls * n_levels * B, dtype=torch.float32, device=dev)
        loss = torch.empty(ncalls * n_levels, dtype=torch.float32, device=dev)

        a = _lib.MtaArgs()
        a.n_levels, a.n_teachers, a.B, a.C = n_levels, n_teachers, B, Cch
        a.dtype = _lib.MMD_F32 if dtype == torch.float32 else _lib.MMD_BF16
        a.layout = layout
        a.T, a.p = T, p
        a.separate = 1 if separate else 0
        for l in range(n_levels):
            a.H[l], a.W[l] = fs[l].shape[2], fs[l].shape[3]
            a.fs[l] = fs[l].data_ptr()
            for k in range(n_teachers):
                a.ft[k][l] = feats[(k + 1) * n_levels + l].data_ptr()
        a.att_ws, a.loss_b, a.loss = att.data_ptr(), loss_b.data_ptr(), loss.data_ptr()
        a.ga_ws = ga.data_ptr() if ga is not None else None
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().mmd_mta_fwd(C.byref(a), stream), "mmd_mta_fwd")
        ctx.args = a
        ctx.n_levels = n_levels
        ctx.keep = (fs, ga)          # keeps the pointers inside `a` alive until backward
        ctx.dev = dev
        # always fp32: a bf16 loss could not even represent loss + ln(HW)
        return loss.view(n_teachers, n_levels) if separate else loss

    @staticmethod
    def backward(ctx, grad_loss):
        fs, ga = ctx.keep
        n = ctx.n_levels
        if ga is None:
            return (None,) * len(ctx.needs_input_grad)
        go = grad_loss.detach().to(torch.float32).contiguous()
        grads = [torch.empty_like(f) for f in fs]
        ptrs = (C.c_void_p * n)(*[g.data_ptr() for g in grads])
        with torch.cuda.device(ctx.dev):
            stream = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().mmd_mta_bwd(C.byref(ctx.args), go.data_ptr(), ptrs, stream), "mmd_mta_bwd")
        out = [None, None, None, None, None]
        for l in range(n):
            out.append(grads[l] if ctx.needs_input_grad[5 + l] else None)
        out.extend([None] * (len(ctx.needs_input_grad) - len(out)))
        return tuple(out)


class MTALoss(nn.Module):
    """Multi-teacher alignment loss; signature and semantics of src/loss/MTALoss.py:9-77."""

    def __init__(self, T=9.0, p=2.0):
        super(MTALoss, self).__init__()
        self.p = float(p)   # src/loss/MTALoss.py:12-13 (config passes strings)
        self.T = float(T)

    def _run(self, g_s, teachers, separate=False):
        n_levels = len(g_s)
        if not 1 <= n_levels <= _lib.MTA_MAX_LEVELS:
            raise ValueError("MTALoss supports 1..%d pyramid levels, got %d" % (_lib.MTA_MAX_LEVELS, n_levels))
        if not 1 <= len(teachers) <= _lib.MTA_MAX_TEACHERS:
            raise ValueError("MTALoss supports 1..%d teachers per call, got %d" % (_lib.MTA_MAX_TEACHERS, len(teachers)))
        flat = list(g_s)
        for t in teachers:
            flat.extend(t[:n_levels])
        return _MTAFunction.apply(self.T, self.p, n_levels, len(teachers), bool(separate), *flat)

    def forward(self, g_s, g_t):
        if torch.is_tensor(g_t[0]):                     # one teacher: list of level tensors (MTALoss.py:17-19)
            n = min(len(g_s), len(g_t))                 # zip() semantics
            return self._run(list(g_s)[:n], [list(g_t)[:n]])
        n = len(g_s)                                    # list of teachers, each a list of levels (:20-34)
        return self._run(list(g_s), [list(t)[:n] for t in g_t])

    def forward_each(self, g_s, teachers):
        """`torch.stack([self(g_s, t) for t in teachers])` — the per-teacher calls of the reference's step wrappers
        (train_methods.py:351-358) — in ONE set of launches: the student maps are pooled once, the backward writes the
        sum of the calls' gradients in one pass.  Returns Tensor[len(teachers), len(g_s)]."""
        n = len(g_s)
        return self._run(list(g_s), [list(t)[:n] for t in teachers], separate=True)

    def mtaloss(self, out_s, out_t):
        """One level; `out_t` is a tensor or a list of teacher tensors (MTALoss.py:36-74)."""
        teachers = [[out_t]] if torch.is_tensor(out_t) else [[t] for t in out_t]
        return self._run([out_s], teachers)[0]

    def at(self, f):
        """L2-normalised channel-pooled attention map [B, H*W] (MTALoss.py:76-77); not differentiable here."""
        _check([f])
        f = f.detach()
        layout, (f,) = _layout_of([f])
        B, Cch, H, W = f.shape
        att = torch.empty(2 * B * H * W, dtype=torch.float32, device=f.device)
        loss_b = torch.empty(B, dtype=torch.float32, device=f.device)
        loss = torch.empty(1, dtype=torch.float32, device=f.device)
        a = _lib.MtaArgs()
        a.n_levels, a.n_teachers, a.B, a.C = 1, 1, B, Cch
        a.dtype = _lib.MMD_F32 if f.dtype == torch.float32 else _lib.MMD_BF16
        a.layout, a.T, a.p = layout, self.T, self.p
        a.H[0], a.W[0] = H, W
        a.fs[0] = f.data_ptr()
        a.ft[0][0] = f.data_ptr()
        a.att_ws, a.loss_b, a.loss, a.ga_ws = att.data_ptr(), loss_b.data_ptr(), loss.data_ptr(), None
        with torch.cuda.device(f.device):
            _lib.check(_lib.lib().mmd_mta_fwd(C.byref(a), torch.cuda.current_stream().cuda_stream), "mmd_mta_fwd")
        pooled = att[:B * H * W].view(B, H * W)
        return (pooled / pooled.norm(dim=1, keepdim=True).clamp_min(1e-12)).to(f.dtype)
