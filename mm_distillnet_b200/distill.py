"""Data-parallel distillation step over the hot path (one process per GPU).

Mirrors what the reference's step wrapper + driver do around the BiFPN/MTA path
(ModelWithNMSLossAugmented.forward, src/optimization/train_methods.py:310-358: student forward with grad, every
teacher forward under no_grad in eval mode, `criterion_kd(features_s, features_t)` per teacher;
train_traditional, src/optimization/traditional.py:171-182: loss = w_kd * sum(stack(kd_losses)), backward) and
replaces the cfg's DataParallel / DDP gradient exchange (train_methods.py:944-961) by ONE NCCL all-reduce of the
student's flat fp32 gradient buffer (mean over ranks = DDP semantics; BatchNorm statistics stay per rank, MTA's
'batchmean' is over the local batch, teachers are replicated: no other cross-GPU traffic).
"""
import torch
import torch.distributed as dist

from .bifpn import BiFPN, BiFPNStack, forward_multi
from .mta import MTALoss


def _flat_channels_last(like, device):
    """(flat uint8 buffer, views): one channels_last view per tensor of `like` (same logical [B,C,H,W] shape and dtype),
    laid out back to back at 256-byte aligned offsets of ONE allocation."""
    offs, total = [], 0
    for x in like:
        offs.append(total)
        total += (x.numel() * x.element_size() + 255) // 256 * 256
    flat = torch.empty(max(total, 256), dtype=torch.uint8, device=device)
    views = []
    for x, off in zip(like, offs):
        B, C, H, W = x.shape
        v = flat[off:off + x.numel() * x.element_size()].view(x.dtype).view(B, H, W, C).permute(0, 3, 1, 2)
        views.append(v)
    return flat, views


class DistillStep:
    """step(student_inputs, teacher_inputs) -> detached fp32 tensor [n_teachers, n_levels] of MTA losses.

    `student` is a BiFPNStack / BiFPN in train mode, `teachers` a list of BiFPNStack / BiFPN in eval mode.
    Inputs may live on the host (ideally pinned): they are copied to the device inside the call.  After the call
    every student parameter's `.grad` is a view into `self.flat_grad` (already averaged over the process group).
    """

    def __init__(self, student, teachers, criterion=None, w_kd=0.005, process_group=None, device=None,
                 batch_networks=True, kd_mode="each"):
        if not isinstance(student, (BiFPN, BiFPNStack)):
            raise TypeError("DistillStep drives mm_distillnet_b200 BiFPN / BiFPNStack modules")
        self.student, self.teachers = student, list(teachers)
        self.criterion = criterion if criterion is not None else MTALoss()
        self.w_kd = float(w_kd)
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.device = device if device is not None else next(student.parameters()).device
        self.flat_grad = None
        self.batch_networks = bool(batch_networks)
        # "each": criterion_kd(features_s, features_t) per teacher, the shipped ModelWithNMSLossAugmented wrapper
        #         (train_methods.py:351-358) -> losses [n_teachers, n_levels];
        # "product": ONE call criterion_kd(features_s, [features_t1, features_t2, ...]) against the product of the
        #         teachers' attention maps, the ModelWithNMSKDListLoss wrapper (train_methods.py:165-262, KD term :256-261;
        #         MTALoss.forward list-of-lists branch, MTALoss.py:20-34) -> losses [1, n_levels]
        if kd_mode not in ("each", "product"):
            raise ValueError("DistillStep: kd_mode must be 'each' or 'product'")
        self.kd_mode = kd_mode
        # multi-stream fallback (stacks that cannot share launches): the frozen teachers and the student are independent until the MTA loss: each teacher stack runs on its own
        # CUDA stream so the small pyramid levels (P5-P7: fewer CTAs than SMs) of different networks overlap
        self.streams = [torch.cuda.Stream(device=self.device) for _ in self.teachers] if self.device.type == "cuda" else []
        student._runner.grad_sink = self._on_flat_grad
        for t in self.teachers:
            t.eval()
            for p in t.parameters():
                p.requires_grad_(False)     # train_methods.py:891-893

    def _on_flat_grad(self, flat):
        """Called from the student's backward with the single contiguous fp32 gradient buffer."""
        if self.world > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.pg)
            flat.mul_(1.0 / self.world)
        self.flat_grad = flat

    def _to_device(self, xs):
        return tuple(x if x.device == self.device else x.to(self.device, non_blocking=True) for x in xs)

    def _batchable(self, xs, xts):
        stacks = [self.student] + self.teachers
        if len(stacks) > 4 or not all(isinstance(m, BiFPNStack) and m.fusable() for m in stacks):
            return False
        return all(x.dtype == torch.bfloat16 for x in xs) and all(x.dtype == torch.bfloat16 for t in xts for x in t) and \
            all(t[0].shape[0] == xs[0].shape[0] for t in xts)

    def _kd(self, feats_s, feats_t_all):
        """criterion_kd(features_s, features_t) per teacher (train_methods.py:351-358) -> Tensor[n_teachers, n_levels]."""
        teachers = [[f.detach() for f in ft] for ft in feats_t_all]
        if self.kd_mode == "product":
            return self.criterion(feats_s, teachers if len(teachers) > 1 else teachers[0]).unsqueeze(0)
        if isinstance(self.criterion, MTALoss) and 1 <= len(teachers) <= 4:
            return self.criterion.forward_each(feats_s, teachers)      # one set of launches for all teachers
        return torch.stack([self.criterion(feats_s, t) for t in teachers])

    def __call__(self, student_inputs, teacher_inputs):
        xs = self._to_device(student_inputs)
        xts = [self._to_device(t) for t in teacher_inputs]
        if self.batch_networks and self._batchable(xs, xts):
            # one lockstep pass: the same node of the student and of every teacher shares a launch
            outs = forward_multi([(self.student, xs)] + [(t, x) for t, x in zip(self.teachers, xts)])
            feats_s, feats_t_all = outs[0], outs[1:]
            kd = self._kd(feats_s, feats_t_all)                                                  # :351-358
        else:
            main = torch.cuda.current_stream(self.device)
            feats_t_all = []
            for teacher, tin, st in zip(self.teachers, xts, self.streams):   # train_methods.py:320-336
                st.wait_stream(main)
                with torch.cuda.stream(st), torch.no_grad():
                    feats_t = teacher(tin)
                    for f in feats_t:
                        f.record_stream(main)
                feats_t_all.append(feats_t)
            feats_s = self.student(xs)                                   # :318
            for st in self.streams:
                main.wait_stream(st)
            kd = self._kd(feats_s, feats_t_all)                          # :351-358
        # loss = w_kd * sum(stack(kd_losses)); loss.backward()  (traditional.py:171-182, KD term): d loss / d kd[t, l] is
        # the constant w_kd, so the backward starts from that constant instead of three scalar kernels (sum, mul, expand)
        g = getattr(self, "_kd_grad", None)
        if g is None or g.shape != kd.shape or g.device != kd.device:
            g = self._kd_grad = torch.full_like(kd.detach(), self.w_kd)
        torch.autograd.backward([kd], [g])
        return kd.detach()

    # ---- CUDA-graph replay of the whole step -----------------------------------------------------------------------
    def capture(self, student_inputs, teacher_inputs, warmup=3):
        """Capture one full step (teacher forwards on their side streams, student forward, MTA, backward, gradient
        all-reduce) into a CUDA graph over static device copies of the inputs.  The step is ~400 kernel launches of a
        few microseconds each: replaying it as one graph takes the host (and any driver contention, e.g. a clock
        monitor) out of the critical path.  Returns self; use replay()."""
        dev = self.device
        # all static inputs are channels_last views into ONE flat buffer: replay_prefetched() moves a whole staging set
        # into them with a single device-to-device copy
        srcs = list(student_inputs) + [x for xs in teacher_inputs for x in xs]
        self._flat_static, views = _flat_channels_last(srcs, dev)
        with torch.no_grad():
            for v, x in zip(views, srcs):
                v.copy_(x.detach())
        ns = len(student_inputs)
        self._g_xs = [v.requires_grad_(bool(x.requires_grad)) for v, x in zip(views[:ns], student_inputs)]
        self._g_xt, k = [], ns
        for xs in teacher_inputs:
            self._g_xt.append(views[k:k + len(xs)])
            k += len(xs)
        self._stage_xs = None
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(int(warmup), 1)):      # plans, arenas and packed blocks exist before the capture starts
                for x in self._g_xs:
                    x.grad = None
                self(self._g_xs, self._g_xt)
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        for x in self._g_xs:
            x.grad = None
        try:   # the static inputs' AccumulateGrad nodes were created on another stream than the capture stream
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        except AttributeError:
            pass
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph, stream=side):   # the stream the warm-up ran on (autograd leaf streams match)
            self._g_out = self(self._g_xs, self._g_xt)
        return self

    def replay(self, student_inputs=None, teacher_inputs=None):
        """Copy new inputs (host or device tensors; pinned host memory makes the copies asynchronous) into the static
        buffers and replay the captured step.  Returns the static [n_teachers, n_levels] loss tensor; the student's
        parameter gradients are in `flat_grad` / `.grad`, input gradients in `graph_inputs()[i].grad`."""
        if getattr(self, "_graph", None) is None:
            raise RuntimeError("DistillStep.replay() needs capture() first")
        if student_inputs is not None:
            for d, x in zip(self._g_xs, student_inputs):
                d.detach().copy_(x, non_blocking=True)
        if teacher_inputs is not None:
            for ds, xs in zip(self._g_xt, teacher_inputs):
                for d, x in zip(ds, xs):
                    d.copy_(x, non_blocking=True)
        self._graph.replay()
        return self._g_out

    # ---- pipelined input feed for replay(): the host->device copy of step i+1 overlaps the replay of step i -----------
    def prefetch(self, student_inputs, teacher_inputs):
        """Start copying the NEXT step's inputs (host tensors, ideally pinned) into a device staging set on a side
        stream.  The copy waits until the previous replay_prefetched() has consumed the staging set, so one call per
        step is safe; it never blocks the host for pinned inputs."""
        if getattr(self, "_graph", None) is None:
            raise RuntimeError("DistillStep.prefetch() needs capture() first")
        dev = self.device
        if getattr(self, "_stage_xs", None) is None:
            self._flat_stage, views = _flat_channels_last([x.detach() for x in self._g_xs] + [x for xs in self._g_xt for x in xs], dev)
            ns = len(self._g_xs)
            self._stage_xs, self._stage_xt, k = views[:ns], [], ns
            for xs in self._g_xt:
                self._stage_xt.append(views[k:k + len(xs)])
                k += len(xs)
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._ev_ready = torch.cuda.Event()
            self._ev_consumed = torch.cuda.Event()
            self._ev_consumed.record(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._ev_consumed)
            for d, x in zip(self._stage_xs, student_inputs):
                d.copy_(x, non_blocking=True)
            for ds, xs in zip(self._stage_xt, teacher_inputs):
                for d, x in zip(ds, xs):
                    d.copy_(x, non_blocking=True)
            self._ev_ready.record(self._copy_stream)
        self._prefetched = True

    def replay_prefetched(self):
        """Replay the captured step on the inputs of the last prefetch(): waits for that copy, moves the staging set into
        the graph's static inputs (device-to-device, ~0.1 ms for the D2 pyramid at B=16) and replays."""
        if not getattr(self, "_prefetched", False):
            raise RuntimeError("DistillStep.replay_prefetched() needs prefetch() first")
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self._ev_ready)
        with torch.no_grad():
            self._flat_static.copy_(self._flat_stage)       # one copy: both sets share the flat layout
        self._ev_consumed.record(cur)
        self._prefetched = False
        self._graph.replay()
        return self._g_out

    def graph_inputs(self):
        return self._g_xs
