"""One-pass GPU diagnostics: every parity case, every metric, nothing asserted.  Writes gpurun_out/diag.json.
    python tests/run_gpu_diag.py [filter]
"""
import json
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from tests import gpu_cases as G  # noqa: E402

BF = torch.bfloat16
CASES = [
    ("mta_golden_c112_nchw", lambda: G.mta_golden_case("mta_c112")),
    ("mta_golden_c112_nhwc", lambda: G.mta_golden_case("mta_c112", channels_last=True)),
    ("mta_golden_c16_nhwc", lambda: G.mta_golden_case("mta_c16", channels_last=True)),
    ("mta_rand_t1", lambda: G.mta_random_case(4, 112, [24, 12, 6, 3, 2], 1)),
    ("mta_rand_t3", lambda: G.mta_random_case(4, 112, [24, 12, 6, 3, 2], 3)),
    ("mta_rand_t3_uniform", lambda: G.mta_random_case(2, 112, [96, 48, 24, 12, 6], 3, structured=False)),
    ("mta_rand_t1_nchw", lambda: G.mta_random_case(3, 112, [12, 6], 1, channels_last=False)),
    ("mta_rand_t2_bf16", lambda: G.mta_random_case(4, 112, [24, 12, 6], 2, dtype=BF)),
    ("mta_rand_c160_p3", lambda: G.mta_random_case(2, 160, [10, 5], 1, p=3.0)),
    ("stack_golden_cell_c112", lambda: G.stack_golden_case("cell_c112")),
    ("stack_golden_stack2_c112", lambda: G.stack_golden_case("stack2_c112")),
    ("stack_rand_cell_s32", lambda: G.random_stack_case(1, False, 2, 32)),
    ("stack_rand_first_s32", lambda: G.random_stack_case(1, True, 2, 32)),
    ("stack_rand_3cells_s48_mixedfw", lambda: G.random_stack_case(3, True, 2, 48, fw_mode="mixed", channels_last=True)),
    ("stack_rand_5cells_s96", lambda: G.random_stack_case(5, True, 2, 96)),
    ("stack_rand_2cells_s48_oddP7", lambda: G.random_stack_case(2, True, 1, 48)),  # 48,24,12,6,3: odd P7
    ("stack_rand_3cells_s48_bf16", lambda: G.random_stack_case(3, True, 2, 48, dtype=BF)),
    ("stack_rand_5cells_s96_bf16", lambda: G.random_stack_case(5, True, 2, 96, dtype=BF)),
    ("stack_rand_2cells_s48_bf16_smoke", lambda: G.random_stack_case(2, True, 2, 48, dtype=BF, ref64=False)),
    ("stack_golden_cell_c112_bf16", lambda: G.stack_golden_case("cell_c112", dtype=BF)),
    ("stack_golden_stack2_c112_bf16", lambda: G.stack_golden_case("stack2_c112", dtype=BF)),
    # the same fp32 cases WITHOUT the arg-max hints (what the max-pool flips cost), and repeated with hints (stability)
    ("stack_rand_5cells_s96_nohint", lambda: G.random_stack_case(5, True, 2, 96, force_argmax=False)),
    ("stack_rand_3cells_s48_mixedfw_nohint", lambda: G.random_stack_case(3, True, 2, 48, fw_mode="mixed", channels_last=True, force_argmax=False)),
    ("stack_rand_5cells_s96_rep2", lambda: G.random_stack_case(5, True, 2, 96)),
    ("stack_rand_5cells_s96_seed1", lambda: G.random_stack_case(5, True, 2, 96, seed=1)),
    ("stack_rand_3cells_s48_mixedfw_seed1", lambda: G.random_stack_case(3, True, 2, 48, fw_mode="mixed", channels_last=True, seed=1)),
]


def main():
    flt = sys.argv[1] if len(sys.argv) > 1 else ""
    out = {}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for name, fn in CASES:
        if flt and flt not in name:
            continue
        t0 = time.time()
        try:
            out[name] = fn()
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            out[name] = {"ERROR": "%s: %s" % (type(e).__name__, e), "trace": traceback.format_exc()[-1500:]}
        out[name]["_seconds"] = round(time.time() - t0, 2)
        print(name, json.dumps(out[name], default=float)[:1200], flush=True)
        with open(os.path.join(ROOT, "gpurun_out", "diag.json"), "w") as f:
            json.dump(out, f, indent=1, default=float)


if __name__ == "__main__":
    main()
