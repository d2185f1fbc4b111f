import sys; sys.path.insert(0, "/root/repo")
import torch
from tests import gpu_cases as G
for it in range(3):
    m = G.random_stack_case(5, True, 2, 96, dtype=torch.bfloat16)
    print({k: round(v, 4) for k, v in m.items() if isinstance(v, float) and (k.startswith("train_") or k.startswith("torchbf16_train") or k.startswith("grad_in") or k.startswith("torchbf16_grad") or k.startswith("eval_"))})
