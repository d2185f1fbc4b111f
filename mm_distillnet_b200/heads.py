"""Drop-ins for the reference's detection heads (src/YetAnotherEfficientDet.py:445-532) on the sm_100a BiFPN kernels.

`Regressor` / `Classifier` keep the reference constructor and forward signatures and the exact `state_dict` names
(conv_list.<i>.{depthwise_conv,pointwise_conv}.conv.*, bn_list.<level>.<i>.*, header.*), so reference checkpoints load
with strict `load_state_dict`.  A forward is ONE op list through `mmd_bifpn_run` (see `_Plan._head` in bifpn.py): the
conv towers are one-input fusion nodes whose BatchNorm + swish is applied by the next kernel while it loads, the header
convolution runs on zero-padded weights, and a gather kernel writes the concatenated `[B, sum HW * anchors, k]` result.

CUDA only, C = 112 (EfficientDet-D2); parameters stay fp32, activations are float32 or bfloat16 NHWC.  No CPU fallback.
"""
import torch.nn as nn

from .bifpn import BN_EPS, BN_MOMENTUM, KERNEL_CHANNELS, SeparableConvBlock, _Runner, _Stateless


class _Head(nn.Module):
    sigmoid_output = False

    def __init__(self, in_channels, num_anchors, out_per_anchor, num_layers, onnx_export=False):
        super().__init__()
        if in_channels != KERNEL_CHANNELS:
            raise NotImplementedError("mm_distillnet_b200.%s: the sm_100a kernels are built for in_channels=%d "
                                      "(EfficientDet-D2), got %d" % (type(self).__name__, KERNEL_CHANNELS, in_channels))
        if num_layers < 1 or num_anchors * out_per_anchor > 2 * KERNEL_CHANNELS:
            raise NotImplementedError("mm_distillnet_b200.%s: num_layers >= 1 and at most %d header channels are supported"
                                      % (type(self).__name__, 2 * KERNEL_CHANNELS))
        self.in_channels = in_channels
        self.num_anchors = num_anchors
        self.num_layers = num_layers
        # construction order mirrors the reference so the same torch seed gives the same initial weights
        self.conv_list = nn.ModuleList(
            [SeparableConvBlock(in_channels, in_channels, norm=False, activation=False) for i in range(num_layers)])
        self.bn_list = nn.ModuleList(
            [nn.ModuleList([nn.BatchNorm2d(in_channels, momentum=BN_MOMENTUM, eps=BN_EPS) for i in range(num_layers)])
             for j in range(5)])
        self.header = SeparableConvBlock(in_channels, num_anchors * out_per_anchor, norm=False, activation=False)
        self.swish = _Stateless()
        self._runner = _Runner()

    def forward(self, inputs):
        """inputs: the 5 pyramid levels (P3..P7).  Returns (predictions, alignment) like the reference: the predictions of
        every level concatenated along dim 1, and the activated last tower layer of the LAST level."""
        out, align = self._runner.run([self], tuple(inputs), self.training, kind="head")
        return out, align


class Regressor(_Head):
    """Box regression head: -> ([B, sum_l H_l W_l * num_anchors, 4], alignment).  src/YetAnotherEfficientDet.py:445-487."""

    def __init__(self, in_channels, num_anchors, num_layers, onnx_export=False):
        super().__init__(in_channels, num_anchors, 4, num_layers, onnx_export)


class Classifier(_Head):
    """Classification head: -> (sigmoid scores [B, sum_l H_l W_l * num_anchors, num_classes], alignment).
    src/YetAnotherEfficientDet.py:490-533."""
    sigmoid_output = True

    def __init__(self, in_channels, num_anchors, num_classes, num_layers, onnx_export=False):
        super().__init__(in_channels, num_anchors, num_classes, num_layers, onnx_export)
        self.num_classes = num_classes
