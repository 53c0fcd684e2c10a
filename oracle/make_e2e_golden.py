"""Golden results of the UNMODIFIED reference pipeline (build container only, CPU):

    python oracle/make_e2e_golden.py [case ...]       # -> tests/golden/e2e_<case>.json

Runs `python -m src.train --cfg X` then `python -m src.infer --cfg X` with PYTHONPATH=/root/reference ONLY (no overlay,
none of this repo's code on the model path) on the synthetic App.-D cache of tests/e2e_common.py and records the
train_log.csv rows, the submission and the checkpoint layout.  tests/test_gpu_dropin_e2e.py repeats the same two commands
on the B200 box with the b200ctr overlay in front and compares.  TEST INFRASTRUCTURE ONLY.
"""
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import e2e_common as E  # noqa: E402

REF = os.environ.get("B200CTR_REFERENCE", "/root/reference")

if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("reference not present: golden e2e results can only be regenerated in the build container")
    for case in (sys.argv[1:] or list(E.CASES)):
        work = tempfile.mkdtemp(prefix=f"e2e_ref_{case}_")
        try:
            cfg_path = E.prepare(case, work, REF, device="cpu")
            t0 = time.time()
            E.run_module("src.train", cfg_path, [REF], 3600, os.path.join(work, "train.log"))
            t1 = time.time()
            E.run_module("src.infer", cfg_path, [REF], 3600, os.path.join(work, "infer.log"))
            res = E.collect(case, work)
            res["reference_cpu_seconds"] = {"train": t1 - t0, "infer": time.time() - t1, "cores": os.cpu_count()}
            with open(E.golden_path(case), "w") as fh:
                json.dump(res, fh)
            print(case, "train", round(t1 - t0, 1), "s; log:", res["train_log"])
        finally:
            shutil.rmtree(work, ignore_errors=True)
