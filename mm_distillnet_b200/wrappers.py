"""Drop-ins for the reference's step wrappers: the `nn.Module`s that `train_methods.train()` builds around
the student, the frozen teachers and the criteria (src/optimization/train_methods.py:894-918), with the constructor,
`forward(rgb, thermal, depth, audio, label, validate=False, augment=False)` and the 6-entry return value of

    ModelWithNMSLoss            :425-516   per-teacher criterion_kd(features_s, features_t) calls
    ModelWithNMSLossAugmented   :265-422   the same step + the augment = True branch (samples 0 and 1 merged)
    ModelWithNMSKDListLoss      :165-262   ONE criterion_kd(features_s, [features_t ...]) call (multi-teacher product)
    ModelWithNMSKDListLossAugmented :50-162  the same + (augment = True) the rgb teacher on the images in `label` as one more teacher

What changes is where the work between the model outputs and the losses runs.  The reference does, per teacher and per
sample, a `.cpu()` of the over-threshold boxes, torchvision NMS, numpy concatenation, one more NMS on the host and a
host→device copy of the labels inside the detection loss: 3·B host synchronisations per step.  Here the teachers'
predictions go through ONE `mmd_pseudo_labels` call (pseudo.py) whose padded label tensor `YetAnotherFocalLoss` (focal.py)
reads on the device, and the per-teacher KD calls go through `MTALoss.forward_each` (one set of launches): no host
synchronisation between the models' outputs and the loss values.  When the student and its (<= 3, frozen) teachers are
YetAnotherEfficientDet models built from this package's stack and heads (patch_reference(heads=True)), their forwards behind
the backbones also run in lockstep (`lockstep_detection_forward`: same nodes of all networks share launches).

`criterion_main` must be mm_distillnet_b200.YetAnotherFocalLoss (or any callable that accepts `PseudoLabels`; a foreign
criterion gets the reference's list-of-arrays via `.to_list()`, which synchronises); `criterion_kd` any callable with the
reference's signature — mm_distillnet_b200.MTALoss takes the fused path.
"""
import contextlib

import torch
import torch.nn as nn

from .focal import YetAnotherFocalLoss
from .mta import MTALoss
from .pseudo import DEFAULT_CAP, DEFAULT_MAX_LABELS, DEFAULT_MAX_ROWS, teacher_pseudo_labels

_MODALITIES = ("rgb", "audio", "thermal", "depth")


class _NMSStep(nn.Module):
    """Shared body of the three wrappers (their forward methods differ only in how criterion_kd is called)."""

    kd_list = False          # True: one criterion_kd call on the list of teachers (ModelWithNMSKDListLoss[Augmented])
    kd_list_augmented = False   # True: `augment=True` adds the rgb teacher's view of `label` as one more teacher

    def __init__(self, student_model, teacher_models, criterion_main, criterion_div, criterion_kd, config, valid_classes_dict):
        super().__init__()
        self.criterion_main = criterion_main
        self.criterion_div = criterion_div
        self.criterion_kd = criterion_kd
        self.student_model = student_model
        self.teacher_models = teacher_models
        self.config = config
        self.valid_classes_dict = valid_classes_dict
        # capacities of the device-side label generation (the reference's Python lists have none)
        self.pseudo_cap, self.pseudo_max_rows, self.pseudo_max_labels = DEFAULT_CAP, DEFAULT_MAX_ROWS, DEFAULT_MAX_LABELS
        self.last_pseudo_labels = None     # the PseudoLabels of the last call (device-resident; for logging / tests)

    # ---- lockstep across networks ---------------------------------------------------------------------------------------
    lockstep = True          # use lockstep_detection_forward when every model allows it (see _lockstep_models)

    def _lockstep_models(self, extra):
        """The student and its teachers as YetAnotherEfficientDet-shaped modules (`backbone_net`, `bifpn`, `regressor`,
        `classifier`, `anchors`; features_from == 'efficientnet', src/YetAnotherEfficientDet.py:656-680) built from this
        package's BiFPNStack / Regressor / Classifier — or None when anything differs (another model class, the reference's
        own PyTorch heads, more than 3 teachers, a teacher that is not frozen, the `augmentation` teacher), in which case the
        wrappers call one model after the other exactly like the reference."""
        from .bifpn import BiFPNStack
        from .heads import _Head
        if not self.lockstep or extra is not None:
            return None
        models = [self.student_model] + [self.teacher_models[m] for m in self.teacher_models.keys()]
        if not 2 <= len(models) <= 4:
            return None
        for i, m in enumerate(models):
            if not all(hasattr(m, a) for a in ("backbone_net", "bifpn", "regressor", "classifier", "anchors")):
                return None
            if getattr(m, "features_from", "efficientnet") != "efficientnet":
                return None
            if not isinstance(m.bifpn, BiFPNStack) or not m.bifpn.fusable() or not isinstance(m.regressor, _Head) or \
                    not isinstance(m.classifier, _Head):
                return None
            if i > 0 and (m.training or any(p.requires_grad for p in m.parameters())):
                return None
        return models

    def _lockstep_outputs(self, models, inputs):
        """YetAnotherEfficientDet.forward (:660-680) for all networks: every backbone on its own input (PyTorch), then the
        stacks, regressors and classifiers in lockstep.  -> logits_s, features_s, predictions, features."""
        from .distill import lockstep_detection_forward
        feats = []
        for i, (m, x) in enumerate(zip(models, inputs)):
            if i == 0:
                _, p3, p4, p5 = m.backbone_net(x)
            else:
                with torch.no_grad():
                    _, p3, p4, p5 = m.backbone_net(x)
            feats.append((p3, p4, p5))
        if len({(f[0].dtype, f[0].shape[0]) for f in feats}) == 1 and feats[0][0].dtype == torch.bfloat16:
            outs = lockstep_detection_forward(models[0], models[1:], feats[0], [tuple(t.detach() for t in f) for f in feats[1:]])
        else:       # fp32 features (the parity mode) or unequal batches: network by network, from the features computed above
            outs = []
            for i, (m, f) in enumerate(zip(models, feats)):
                with contextlib.nullcontext() if i == 0 else torch.no_grad():
                    fe = m.bifpn(f)
                    r, _ = m.regressor(fe)
                    c, _ = m.classifier(fe)
                outs.append((c, r, fe))
        res = []
        for m, x, (c, r, f) in zip(models, inputs, outs):
            res.append(([c, r, m.anchors(x, x.dtype)], f))
        (logits_s, features_s), rest = res[0], res[1:]
        return logits_s, features_s, [p for p, _ in rest], [[t.detach() for t in f] for _, f in rest]

    def _teacher_outputs(self, rgb, thermal, depth, audio, extra=None):
        inputs = {"rgb": rgb, "audio": audio, "thermal": thermal, "depth": depth}
        predictions, features = [], []
        runs = [(m, self.teacher_models[m]) for m in self.teacher_models.keys()]
        if extra is not None:
            # ModelWithNMSKDListLossAugmented (:72-75, :90-95): one more "teacher" = the rgb teacher on the images in `label`
            runs.append(("augmentation", self.teacher_models["rgb"]))
            inputs["augmentation"] = extra
        for modality, teacher_model in runs:
            if modality not in inputs:
                raise ValueError('No valid modality to predict from teacher')          # train_methods.py:453-454
            with torch.no_grad():
                prediction, features_t = teacher_model(inputs[modality])
            if isinstance(features_t, (tuple, list)):
                features_t = [f.detach() for f in features_t]
            else:
                features_t = features_t.detach()
            predictions.append(prediction)
            features.append(features_t)
        return predictions, features

    augmented = False        # True: `augment=True` merges samples 0 and 1 (ModelWithNMSLossAugmented); the other wrappers
                             # accept the argument and ignore it, as the reference's do (:176, :436)

    @staticmethod
    def merge_batch_0_1(audio):
        """train_methods.py:290-308, as written (in place, no grad): audio[1] <- log10(max(audio[0]^10 + audio[1]^10, 1e-7))."""
        with torch.no_grad():
            audio[1] = torch.pow(audio[0], 10) + torch.pow(audio[1], 10)
            eps = 1e-7
            audio[1][audio[1] < eps] = eps
            audio[1] = torch.log10(audio[1])
        return audio

    @staticmethod
    def average_batch_0_1(features_t):
        """train_methods.py:276-288: sample 1 of every teacher feature map <- mean of samples 0 and 1 (in place)."""
        with torch.no_grad():
            for i in range(len(features_t)):
                features_t[i][1] = (features_t[i][0] + features_t[i][1]) / 2
        return features_t

    def forward(self, rgb, thermal, depth, audio, label, validate=False, augment=False):
        extra = label if (augment and self.kd_list_augmented) else None
        augment = bool(augment) and self.augmented
        if augment:
            if rgb.shape[0] < 2:
                raise ValueError("augment=True merges samples 0 and 1: the batch needs at least 2 samples")
            audio = self.merge_batch_0_1(audio)                                         # :315-316
        done = None
        models = self._lockstep_models(extra)
        if models is not None:
            by_name = {"rgb": rgb, "audio": audio, "thermal": thermal, "depth": depth}
            names = list(self.teacher_models.keys())
            if all(n in by_name for n in names):
                done = self._lockstep_outputs(models, [audio] + [by_name[n] for n in names])
        if done is not None:
            logits_s, features_s, predictions, features = done
        else:
            logits_s, features_s = self.student_model(audio)
            predictions, features = self._teacher_outputs(rgb, thermal, depth, audio, extra)
        if augment:
            features = [self.average_batch_0_1(list(f)) if isinstance(f, (list, tuple)) else f for f in features]   # :340-341
        dev = rgb.device
        if len(predictions) > 0:
            with torch.no_grad():
                labels = teacher_pseudo_labels(predictions, self.valid_classes_dict, self.config, cap=self.pseudo_cap,
                                               max_rows=self.pseudo_max_rows, max_labels=self.pseudo_max_labels,
                                               merge_batch_0_1=augment)                # :384-386
            self.last_pseudo_labels = labels
            annotations = labels if isinstance(self.criterion_main, YetAnotherFocalLoss) else labels.to_list()
        else:
            annotations = [[] for _ in range(rgb.shape[0])]
        loss_regression, loss_cls = self.criterion_main(logits_s, annotations)

        if self.kd_list:
            loss_kd = torch.zeros(1)
            if self.criterion_kd is not None:
                loss_kd = self.criterion_kd(features_s, features)
            kd_losses = [loss_kd]
        elif self.criterion_kd is None:
            kd_losses = [torch.zeros(1) for _ in features]
        elif isinstance(self.criterion_kd, MTALoss) and 1 < len(features) <= 4 and \
                all(isinstance(f, (list, tuple)) for f in features):
            kd_losses = list(self.criterion_kd.forward_each(features_s, features).unbind(0))
        else:
            kd_losses = [self.criterion_kd(features_s, f) for f in features]
        z = torch.zeros(1, device=dev)
        return [[loss_regression], [loss_cls], kd_losses, z, z.clone(), z.clone()]


class ModelWithNMSLoss(_NMSStep):
    """src/optimization/train_methods.py:425-516."""


class ModelWithNMSLossAugmented(_NMSStep):
    """src/optimization/train_methods.py:265-422, including the augment=True branch (spectrogram merge of samples 0 and 1,
    averaged teacher features, merged labels)."""
    augmented = True


class ModelWithNMSKDListLoss(_NMSStep):
    """src/optimization/train_methods.py:165-262."""
    kd_list = True


class ModelWithNMSKDListLossAugmented(_NMSStep):
    """src/optimization/train_methods.py:50-162: the list-loss step; with augment=True the rgb teacher also sees the images
    passed as `label`, and its predictions / features join the label integration and the KD list as one more teacher."""
    kd_list = True
    kd_list_augmented = True
