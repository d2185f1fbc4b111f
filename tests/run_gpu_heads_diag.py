"""TEST INFRASTRUCTURE (imports the oracle).  Dump every parity metric of the detection heads (fp32 and bf16 storage) plus an all-bf16 PyTorch run of the oracle on the
same GPU as the calibration for the bf16 bounds.  Writes gpurun_out/r2_heads_diag.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mmd_oracle as O   # noqa: E402
from tests import helpers as H   # noqa: E402
from tests import test_gpu_heads as T   # noqa: E402


torch_bf16 = T.torch_bf16


def main():
    out = {}
    for name in sorted(H.HEAD_CASES):
        kind, C, A, K, L, B, s3, seed = H.HEAD_CASES[name]
        params, xs = H.head_case_inputs(name)
        g = H.golden(name)
        gf = lambda y, a: H.head_case_gouts(name, y, a)   # noqa: E731
        for dt, tag in ((torch.float32, "f32"), (torch.bfloat16, "bf16")):
            out["%s_%s" % (name, tag)] = T.head_metrics(T._run_case(kind, C, A, K, L, params, xs, gf, dt), g)
        out["%s_torchbf16" % name] = T.head_metrics(torch_bf16(kind, C, A, K, L, params, xs, gf), g)
    for kind, K, s3, B in (("reg", 20, 64, 2), ("cls", 20, 64, 2), ("cls", 3, 64, 2), ("cls", 20, 96, 4)):
        C, A, L, seed = 112, 9, 3, 31
        params = O.synth_head_params(C, A * (4 if kind == "reg" else K), L, seed)
        xs = H.pyramid_inputs(B, C, s3, seed + 50)
        gf = lambda y, a: (O.synth(tuple(y.shape), seed + 70, 1.0, 0.0), O.synth(tuple(a.shape), seed + 71, 1.0, 0.0))   # noqa: E731
        g = T.oracle_reference(kind, C, A, K, L, params, xs, gf)
        tag = "%s_K%d_s%d_B%d" % (kind, K, s3, B)
        for dt, t2 in ((torch.float32, "f32"), (torch.bfloat16, "bf16")):
            out["%s_%s" % (tag, t2)] = T.head_metrics(T._run_case(kind, C, A, K, L, params, xs, gf, dt), g)
        out["%s_torchbf16" % tag] = T.head_metrics(torch_bf16(kind, C, A, K, L, params, xs, gf), g)
        gq = T.quad_gouts(seed)
        g = T.oracle_reference(kind, C, A, K, L, params, xs, gq)
        out["%s_quad_bf16" % tag] = T.head_metrics(T._run_case(kind, C, A, K, L, params, xs, gq, torch.bfloat16), g)
        out["%s_quad_torchbf16" % tag] = T.head_metrics(torch_bf16(kind, C, A, K, L, params, xs, gq), g)
    for kind, K, s3, B in (("reg", 20, 64, 2), ("cls", 20, 64, 2), ("cls", 3, 64, 2), ("cls", 20, 96, 4)):
        C, A, L, seed = 112, 9, 3, 77
        params, xs = T.rand_case(C, A * (4 if kind == "reg" else K), L, B, s3, seed)
        tag = "rand_%s_K%d_s%d_B%d" % (kind, K, s3, B)
        for nm, gf in (("synthg", lambda y, a: (O.synth(tuple(y.shape), seed + 70, 1.0, 0.0), O.synth(tuple(a.shape), seed + 71, 1.0, 0.0))),
                       ("quad", T.quad_gouts(seed))):
            g = T.oracle_reference(kind, C, A, K, L, params, xs, gf)
            out["%s_%s_f32" % (tag, nm)] = T.head_metrics(T._run_case(kind, C, A, K, L, params, xs, gf, torch.float32), g)
            out["%s_%s_bf16" % (tag, nm)] = T.head_metrics(T._run_case(kind, C, A, K, L, params, xs, gf, torch.bfloat16), g)
            out["%s_%s_torchbf16" % (tag, nm)] = T.head_metrics(torch_bf16(kind, C, A, K, L, params, xs, gf), g)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/r2_heads_diag.json", "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    for k, v in out.items():
        print(k, {kk: round(vv, 6) for kk, vv in v.items() if kk in ("fwd", "grad_in", "pgrad", "buf", "train_align", "train_out", "eval_out", "eval_align")})
        worst = sorted(((vv, kk) for kk, vv in v.items() if kk.startswith("pgrad_")), reverse=True)[:4]
        print("   worst pgrad:", [(kk, round(vv, 4)) for vv, kk in worst])
        print("   grad_in:", [(kk, round(vv, 4)) for kk, vv in sorted(v.items()) if kk.startswith("grad_in")])


if __name__ == "__main__":
    main()
