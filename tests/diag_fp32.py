import sys; sys.path.insert(0, "/root/repo")
import torch
from tests import gpu_cases as G
for it in range(4):
    m = G.random_stack_case(5, True, 2, 96)
    print({k: (round(v, 5) if isinstance(v, float) else v) for k, v in m.items() if "pgrad" in k or "worstname" in k or "grad_in" in k})
