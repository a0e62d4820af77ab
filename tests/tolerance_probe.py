#!/usr/bin/env python
"""How far a build of the library with relaxed arithmetic (FMA contraction, ...) is from the reference: relative L1
per variable on every golden case (tests/golden/, outputs of the unmodified reference) — after one step against the
oracle (north-star bar 1e-12) and after all steps of the case against the reference's grid (bar 1e-10) — plus the
dt sequence.  Test infrastructure (imports the oracle); run on the GPU box:

    ASTREA_B200_LIB=astrea_b200/lib/variants/fma.so python tests/tolerance_probe.py > gpurun_out/tolerance_fma.json
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

from conftest import golden_case, golden_index, rel_l1  # noqa: E402
from cases import run_native, run_oracle  # noqa: E402


def main():
    from astrea_b200 import _native
    lib = _native.device_library()
    rows = {}
    for cid in sorted(golden_index()):
        meta, data = golden_case(cid)
        row = {}
        try:
            want, dts = run_oracle(meta, data["g0"], 1)
            got, used, _ = run_native(lib, meta, data["g0"], 1)
            row["one_step_rel_l1"] = float(np.max(rel_l1(got, want)))
            row["one_step_dt_rel"] = abs(used[0] - dts[0]) / dts[0]
            got, _, _ = run_native(lib, meta, data["g0"], meta["steps"], dts=list(data["dts"]))
            row["all_steps_rel_l1"] = float(np.max(rel_l1(got, data["g"])))
            row["steps"] = meta["steps"]
            row["bit_identical"] = bool(np.array_equal(got, data["g"], equal_nan=True))
        except Exception as err:      # a non-finite wave speed where the reference had none, ...
            row["error"] = f"{type(err).__name__}: {err}"
        rows[cid] = row
        print(cid, row, file=sys.stderr)
    worst1 = max((r.get("one_step_rel_l1", 0) for r in rows.values()))
    worstn = max((r.get("all_steps_rel_l1", 0) for r in rows.values()))
    print(json.dumps({"library": os.environ.get("ASTREA_B200_LIB", "default"), "worst_one_step": worst1, "worst_all_steps": worstn,
                      "errors": [c for c, r in rows.items() if "error" in r], "cases": rows}, indent=1))


if __name__ == "__main__":
    main()
