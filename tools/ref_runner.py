"""Child process that drives the UNMODIFIED reference package (baseline/_ref/gemlite, Triton kernels) on the GPU box.

    python tools/ref_runner.py JOB.pt OUT.pt

JOB = torch.save'd {"cases": [...], "tensors": {...}, "timing": [...], "autotune": "default"|"fast"|"max"|dict}.
Every case is executed by tests/refdual.py -- the very same code the test process runs against gemlite_b200.
Nothing of this repo's product is imported here (the two packages register the same torch op name)."""
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def main():
    job_path, out_path = sys.argv[1], sys.argv[2]
    if not os.path.isdir(os.path.join(REF, "gemlite")):
        print("baseline/_ref/gemlite is absent", file=sys.stderr)
        sys.exit(3)
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import gemlite                     # needs an NVIDIA driver at import (triton_kernels/utils.py:130)
    import gemlite.helper, gemlite.quant_utils, gemlite.dtypes  # noqa: F401,E401
    import refdual
    assert "baseline/_ref" in gemlite.__file__, gemlite.__file__
    job = torch.load(job_path, weights_only=False)
    gemlite.reset_config()
    gemlite.set_autotune(job.get("autotune", "default"))
    res = {"cases": {}, "timing": {}, "version": gemlite.__version__, "triton": __import__("triton").__version__,
           "gpu": torch.cuda.get_device_name(0)}
    deadline = time.time() + float(job.get("budget_s", 1e9))
    for case in job.get("cases", []):
        if time.time() > deadline:
            res["cases"][case["name"]] = {"skipped": "time budget"}
            continue
        t0 = time.time()
        try:
            res["cases"][case["name"]] = refdual.run_case(gemlite, case, job.get("tensors"))
        except Exception:  # noqa: BLE001
            res["cases"][case["name"]] = {"fatal": traceback.format_exc()[-1500:]}
        res["cases"][case["name"]]["seconds"] = time.time() - t0
        print(f"[ref] {case['name']}: {time.time() - t0:.1f}s", flush=True)
        torch.save(res, out_path)
    for t in job.get("timing", []):
        if time.time() > deadline:
            res["timing"][t["name"]] = {"skipped": "time budget"}
            continue
        if "autotune" in t:
            gemlite.set_autotune(t["autotune"])
        t0 = time.time()
        try:
            res["timing"][t["name"]] = refdual.time_case(gemlite, t)
        except Exception:  # noqa: BLE001
            res["timing"][t["name"]] = {"fatal": traceback.format_exc()[-1500:]}
        res["timing"][t["name"]]["seconds"] = time.time() - t0
        print(f"[ref] timing {t['name']}: {res['timing'][t['name']]}", flush=True)
        torch.save(res, out_path)
    torch.save(res, out_path)


if __name__ == "__main__":
    main()
