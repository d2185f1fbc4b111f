"""One-shot diagnostic: element-wise differences of the dense pseudo-label case (device rows vs oracle rows)."""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mm_distillnet_b200 import pseudo as PS   # noqa: E402
from oracle import mmd_oracle as O   # noqa: E402
from tests import helpers as H   # noqa: E402

size, K, B = 768, 20, 2
anchors = H.efficientdet_anchors(size)
N = anchors.shape[1]
gen = torch.Generator().manual_seed(11)
c = 0.02 + 0.2 * torch.rand(B, N, K, generator=gen)
hot = torch.randperm(N, generator=gen)[:3000]
valid = torch.tensor(H.PSEUDO_VALID_IDS)
c[0, hot, valid[torch.randint(0, len(valid), (3000,), generator=gen)]] = 0.3 + 0.69 * torch.rand(3000, generator=gen)
r = 0.2 * torch.randn(B, N, 4, generator=gen)
ref = O.detections(c, r, anchors, H.PSEUDO_VALID_IDS, image_size=size, **H.PSEUDO_CFG)[0].numpy()
out = PS.teacher_pseudo_labels([(c.cuda(), r.cuda(), anchors.cuda())], H.pseudo_valid_classes_dict(), H.pseudo_config(size),
                               raw_rows=True, max_rows=4096, max_labels=4096)
got = out.teacher_lists()[0][0]
d = got != ref
print("rows", got.shape, "mismatching elements", int(d.sum()), "per column", d.sum(0).tolist())
bad = np.argwhere(d)
for i, j in bad[:8]:
    a, b = got[i, j], ref[i, j]
    print(i, j, repr(a), repr(b), "ulps", int(np.abs(a.view(np.int32) - b.view(np.int32))))
dec_gpu = O.decode_boxes(anchors.cuda(), r.cuda()).cpu()
dec_cpu = O.decode_boxes(anchors, r)
print("torch CUDA decode vs torch CPU decode: differing elements", int((dec_gpu != dec_cpu).sum()), "of", dec_cpu.numel())
