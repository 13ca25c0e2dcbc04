"""Run every kernel check and print a table; never stops at the first failure (one gpurun call
should reveal as much as possible).  usage: python tools/gpu_diag.py [simt|tc] [group ...]"""
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import kernel_checks as kc  # noqa: E402
from variational_mmt_b200 import _lib, ops  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "simt"
groups = set(sys.argv[2:])
ops.set_gemm_mode(1 if mode == "simt" else 0)
scale = 1.0 if mode == "simt" else 100.0
bad = 0
for name, fn in kc.ALL:
    if groups and name not in groups:
        continue
    t0 = time.time()
    try:
        res = fn()
        torch.cuda.synchronize()
    except Exception:
        print(f"[{name}] EXCEPTION\n{traceback.format_exc()}")
        bad += 1
        continue
    for label, err, tol in res:
        ok = err <= tol * scale or (tol == 0.0 and err == 0.0)
        bad += not ok
        print(f"{'ok  ' if ok else 'FAIL'} {label:58s} err={err:.3e} tol={tol * scale:.1e}")
    print(f"[{name}] {time.time() - t0:.1f}s")
print("FAILURES:", bad)
