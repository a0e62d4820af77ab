"""Small runs of every kernel family for `compute-sanitizer --tool memcheck|racecheck python tests/sanitize_smoke.py`
(SURVEY.md §5).  Not a pytest module."""
import sys
sys.path.insert(0, '.')
import numpy as np
from astrea_b200 import _native as N
from astrea_b200.selectors import make_cfg
from astrea_b200.initial import initial_state, problem
for config, cells, dim, sub, sol, ts, mhd, bc in (("ll3", 70, 2, "ppm", "hllc", "ssprk(3,3)", False, None), ("orszag-tang", 66, 2, "plm", "hlld", "ssprk(2,2)", True, None),
                                                   ("sod", 300, 1, "weno5", "lf", "ssprk(2,2)", False, None), ("ll4", 50, 2, "weno7", "lf", "euler", False, "edge"),
                                                   ("orszag-tang", 40, 2, "pcm", "hllc", "ssprk(10,4)", True, "edge")):
    pr = problem(config, cells, 1.4)
    b = bc or pr["boundary"]
    cfg = make_cfg(dimension=dim, cells=cells, boundary=b, gamma=1.4, dx=pr["dx"], cfl=.5, subgrid=sub, solver=sol, timestep=ts, magnetic_2d=mhd)
    ctx = N.Context(cfg)
    ctx.upload(initial_state(config, cells, dim, 1.4, sub in ("ppm", "weno5", "weno7"), boundary=b))
    try:
        ctx.step()
        ctx.set_time(0.0, 0.0)
        ctx.step_async()
        print(config, sub, sol, ctx.get_time(), np.isfinite(ctx.download(primitive=True)).all(), ctx.diagnostics()[0][:2])
    except np.linalg.LinAlgError as err:         # some of the reference's configurations die within a few steps
        print(config, sub, sol, "non-finite:", err)
    ctx.close()
