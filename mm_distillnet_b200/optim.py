"""Optimizer step of the distillation recipe on one sm_100a launch (cfg 3).

`FlatAdam` is `torch.optim.Adam` / `AdamW` as the reference constructs them (src/optimization/train_methods.py:825-842:
`lr`, `betas` from the config, eps 1e-8) for parameters whose gradients already sit in ONE flat fp32 buffer — what
`DistillStep` leaves in `flat_grad` (every parameter's `.grad` is a view of it).  The reference's `optimizer.step()`
(src/optimization/traditional.py:190) walks ~360 small tensors with several ATen kernels each; here all of them are one
`mmd_adam_step` call (one launch + a 1-thread launch that advances the device-side step count), capturable into the step's
CUDA graph.  The parameters stay PyTorch's own fp32 tensors; the moments are two flat buffers with the gradient's layout.
CUDA only; there is no CPU fallback.
"""
import ctypes as C

import torch

from . import _lib

CHUNK = 1024


class FlatAdam(object):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled_weight_decay=False):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatAdam: no trainable parameters")
        if any((not p.is_cuda) or p.dtype != torch.float32 or not p.is_contiguous() for p in self.params):
            raise RuntimeError("FlatAdam needs contiguous float32 CUDA parameters (there is no CPU fallback)")
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.weight_decay, self.decoupled = float(weight_decay), bool(decoupled_weight_decay)
        self.device = self.params[0].device
        self._layout = None          # built from the first flat gradient (offsets of every parameter's .grad view)
        self.step_count = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.exp_avg = self.exp_avg_sq = None

    def _build(self, flat):
        base = flat.storage_offset()
        offs, chunks = [], []
        for i, p in enumerate(self.params):
            g = p.grad
            if g is None or g.untyped_storage().data_ptr() != flat.untyped_storage().data_ptr() or not g.is_contiguous():
                raise RuntimeError("FlatAdam: every parameter's .grad must be a contiguous view of the flat gradient buffer "
                                   "(run one DistillStep first)")
            off = g.storage_offset() - base
            if off < 0 or off + p.numel() > flat.numel():
                raise RuntimeError("FlatAdam: a .grad view lies outside the flat gradient buffer")
            offs.append(off)
            for first in range(0, p.numel(), CHUNK):
                chunks.append((i, first, min(CHUNK, p.numel() - first)))
        dev = self.device
        self._offsets = torch.tensor(offs, dtype=torch.int64, device=dev)
        self._chunks = torch.tensor(chunks, dtype=torch.int64, device=dev).contiguous()
        self._ptrs = torch.tensor([p.data_ptr() for p in self.params], dtype=torch.int64, device=dev)
        self._ptr_key = tuple(p.data_ptr() for p in self.params)
        self.exp_avg = torch.zeros(flat.numel(), dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(flat.numel(), dtype=torch.float32, device=dev)
        self._layout = (tuple(offs), flat.numel())
        self._n = sum(p.numel() for p in self.params)

    def prepare(self, flat_grad):
        """Build the device tables from the layout of `flat_grad` (host work, not capturable): called by step() on first
        use; call it explicitly before capturing a CUDA graph."""
        if self._layout is None:
            self._build(flat_grad)
        return self

    def step(self, flat_grad):
        """One optimizer step from the flat fp32 gradient (same layout as at prepare())."""
        if flat_grad.dtype != torch.float32 or not flat_grad.is_cuda or not flat_grad.is_contiguous():
            raise RuntimeError("FlatAdam.step: the flat gradient must be a contiguous float32 CUDA tensor")
        self.prepare(flat_grad)
        if flat_grad.numel() != self._layout[1]:
            raise RuntimeError("FlatAdam.step: the flat gradient changed its size (%d -> %d)" % (self._layout[1], flat_grad.numel()))
        if tuple(p.data_ptr() for p in self.params) != self._ptr_key:
            raise RuntimeError("FlatAdam.step: a parameter tensor was reallocated since prepare()")
        a = _lib.AdamArgs()
        a.n_chunks, a.decoupled_weight_decay, a.n_elements = self._chunks.shape[0], 1 if self.decoupled else 0, self._n
        a.lr, a.beta1, a.beta2, a.eps, a.weight_decay = self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay
        a.chunks, a.offsets, a.params = self._chunks.data_ptr(), self._offsets.data_ptr(), self._ptrs.data_ptr()
        a.grad, a.exp_avg, a.exp_avg_sq = flat_grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr()
        a.step = self.step_count.data_ptr()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().mmd_adam_step(C.byref(a), torch.cuda.current_stream().cuda_stream), "mmd_adam_step")
        for p in self.params:            # the kernel wrote through raw pointers: bump the tensors' version counters (what
            torch.autograd.graph.increment_version(p)     # an in-place ATen update does; eval-mode plan caches key on them)
        return self
