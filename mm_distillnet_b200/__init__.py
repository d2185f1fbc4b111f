"""mm_distillnet_b200 — B200-native (sm_100a) implementation of MM-DistillNet's distillation hot path:
the EfficientDet-D2 BiFPN stack (forward + backward) and the MTA multi-teacher alignment loss.

Public API mirrors the reference (robot-learning-freiburg/MM-DistillNet):
    BiFPN, SeparableConvBlock      src/YetAnotherEfficientDet.py:154-442
    BiFPNStack                     the nn.Sequential of cells built at src/YetAnotherEfficientDet.py:639-644
    Regressor, Classifier          src/YetAnotherEfficientDet.py:445-533 (detection heads on the same kernels)
    YetAnotherFocalLoss            src/loss/YetAnotherFocalLoss.py:23-190 (detection loss, one launch per direction)
    MTALoss                        src/loss/MTALoss.py:9-77
    logits_to_ground_truth,
    teacher_pseudo_labels          src/utils/utils.py:144-324 + the cross-teacher integration (train_methods.py:360-411):
                                   pseudo-labels made and consumed on the device
    ModelWithNMSLoss, ModelWithNMSKDListLoss,
    ModelWithNMSLossAugmented      the step wrappers src/optimization/train_methods.py:165-262, :265-422, :425-516
    FlatAdam                       torch.optim.Adam / AdamW as built at src/optimization/train_methods.py:825-842, one launch
                                   over DistillStep's flat gradient buffer
    patch_reference()              rebinds the reference's module globals to these classes (drop-in seam)
"""
from .bifpn import BiFPN, BiFPNStack, SeparableConvBlock  # noqa: F401
from .heads import Classifier, Regressor  # noqa: F401
from .focal import YetAnotherFocalLoss  # noqa: F401
from .mta import MTALoss  # noqa: F401
from .pseudo import PseudoLabels, logits_to_ground_truth, teacher_pseudo_labels  # noqa: F401
from .wrappers import (ModelWithNMSKDListLoss, ModelWithNMSKDListLossAugmented, ModelWithNMSLoss,  # noqa: F401
                       ModelWithNMSLossAugmented)
from .patch import patch_reference, fuse_bifpn_stacks  # noqa: F401
from .distill import DistillStep, lockstep_detection_forward  # noqa: F401
from .optim import FlatAdam  # noqa: F401
from ._lib import build, launch_count  # noqa: F401

__all__ = ["BiFPN", "BiFPNStack", "SeparableConvBlock", "Regressor", "Classifier", "YetAnotherFocalLoss", "MTALoss", "PseudoLabels", "logits_to_ground_truth",
           "teacher_pseudo_labels", "ModelWithNMSLoss", "ModelWithNMSKDListLoss", "ModelWithNMSKDListLossAugmented", "ModelWithNMSLossAugmented", "patch_reference", "fuse_bifpn_stacks", "DistillStep", "lockstep_detection_forward", "FlatAdam", "build",
           "launch_count"]
