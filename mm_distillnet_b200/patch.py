"""Drop-in seam for the reference code base: there is no plugin registry, classes are resolved by
module-global name, so `patch_reference()` rebinds those names before any model / criterion is built.

    import mm_distillnet_b200 as mmd
    mmd.patch_reference()          # then run the reference's train.py / evaluate.py logic unchanged

Rebinds:
    src.YetAnotherEfficientDet.BiFPN                     -> mm_distillnet_b200.BiFPN   (looked up at :639-644)
    src.loss.MTALoss.MTALoss, src.utils.utils.MTALoss    -> mm_distillnet_b200.MTALoss (utils.py:50, :1603-1604)
    src.YetAnotherEfficientDet.Regressor / Classifier    -> mm_distillnet_b200.Regressor / Classifier   (heads=True only;
                                                            looked up at :646-652; at most 224 header channels)
    src.optimization.train_methods.ModelWithNMSLoss / ModelWithNMSKDListLoss / ModelWithNMSLossAugmented,
    logits_to_ground_truth (train_methods, utils)        -> wrappers.py / pseudo.py   (step_wrappers=True only)
and wraps YetAnotherEfficientDet.__init__ so `self.bifpn` (an nn.Sequential of cells) becomes a `BiFPNStack` with
identical children and state_dict keys, which runs all cells as one fused op list.
"""
import sys

import torch.nn as nn

from .bifpn import BiFPN, BiFPNStack
from .mta import MTALoss


def fuse_bifpn_stacks(model):
    """Replace every nn.Sequential made only of our BiFPN cells by a BiFPNStack (same children, same keys)."""
    for name, child in list(model.named_children()):
        if isinstance(child, nn.Sequential) and not isinstance(child, BiFPNStack) and len(child) > 0 and \
                all(isinstance(c, BiFPN) for c in child):
            setattr(model, name, BiFPNStack(*list(child)))
        else:
            fuse_bifpn_stacks(child)
    return model


def patch_reference(det_module=None, loss_module=None, utils_module=None, heads=False, detection_loss=False,
                    step_wrappers=False, train_methods_module=None):
    """Rebind the reference's globals.  Modules default to the already-imported `src.*` modules.  `heads=True` also
    rebinds the detection heads (their header is limited to 224 output channels, e.g. 9 anchors x 24 classes);
    `detection_loss=True` rebinds YetAnotherFocalLoss in src.loss.YetAnotherFocalLoss and src.utils.utils (bound by
    `from ... import` at utils.py:53, instantiated at :1581-1582); `step_wrappers=True` rebinds ModelWithNMSLoss /
    ModelWithNMSKDListLoss / ModelWithNMSLossAugmented in src.optimization.train_methods (looked up at :899-906 when
    train() builds the step) and `logits_to_ground_truth` in src.utils.utils and train_methods (bound by `from ... import`
    at train_methods.py:20-29): pseudo-labels are then made and consumed on the device."""
    import importlib
    det = det_module or importlib.import_module("src.YetAnotherEfficientDet")
    loss = loss_module or importlib.import_module("src.loss.MTALoss")
    det.BiFPN = BiFPN
    if heads:
        from .heads import Classifier, Regressor
        det.Regressor, det.Classifier = Regressor, Classifier
    loss.MTALoss = MTALoss
    utils = utils_module or sys.modules.get("src.utils.utils")
    if utils is not None and hasattr(utils, "MTALoss"):
        utils.MTALoss = MTALoss
    if detection_loss:
        from .focal import YetAnotherFocalLoss
        fl = sys.modules.get("src.loss.YetAnotherFocalLoss")
        if fl is not None:
            fl.YetAnotherFocalLoss = YetAnotherFocalLoss
        if utils is not None and hasattr(utils, "YetAnotherFocalLoss"):
            utils.YetAnotherFocalLoss = YetAnotherFocalLoss
    if step_wrappers:
        from . import pseudo, wrappers
        tm = train_methods_module or sys.modules.get("src.optimization.train_methods")
        if tm is not None:
            for name in ("ModelWithNMSLoss", "ModelWithNMSKDListLoss", "ModelWithNMSLossAugmented", "ModelWithNMSKDListLossAugmented"):
                setattr(tm, name, getattr(wrappers, name))
            tm.logits_to_ground_truth = pseudo.logits_to_ground_truth
        if utils is not None and hasattr(utils, "logits_to_ground_truth"):
            utils.logits_to_ground_truth = pseudo.logits_to_ground_truth
    cls = det.YetAnotherEfficientDet
    if not getattr(cls, "_mmd_patched", False):
        orig_init = cls.__init__

        def __init__(self, *args, **kwargs):
            orig_init(self, *args, **kwargs)
            fuse_bifpn_stacks(self)

        cls.__init__ = __init__
        cls._mmd_patched = True
    return det, loss
