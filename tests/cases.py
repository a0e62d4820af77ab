"""Shared helpers of the parity tests: run a case on a native library and on the oracle."""
import numpy as np

from astrea_b200 import _native as N
from astrea_b200.selectors import make_cfg
from oracle import OracleConfig, advance


def oracle_cfg(meta, eigen="closed"):
    return OracleConfig(config=meta["config"], cells=meta["cells"], dimension=meta["dimension"], subgrid=meta["subgrid"],
                        solver=meta["solver"], timestep=meta["timestep"], boundary=meta["boundary"], dx=meta["dx"],
                        gamma=meta["gamma"], cfl=meta["cfl"], magnetic_2d=meta["magnetic_2d"], eigen=eigen,
                        ppm_author=meta.get("ppm_author", "mc"))


def native_cfg(meta, **geometry):
    return make_cfg(dimension=meta["dimension"], cells=meta["cells"], boundary=meta["boundary"], gamma=meta["gamma"],
                    dx=meta["dx"], cfl=meta["cfl"], subgrid=meta["subgrid"], solver=meta["solver"], timestep=meta["timestep"],
                    magnetic_2d=meta["magnetic_2d"], ppm_author=meta.get("ppm_author", "mc"), **geometry)


def run_native(lib, meta, g0, steps, dts=None, **geometry):
    """``steps`` full steps through the C ABI.  With ``dts`` the two-call seam (evolve_space / evolve_time) is used
    with the given time steps, otherwise astrea_step computes dt on its own."""
    ctx = N.Context(native_cfg(meta, **geometry), lib=lib)
    try:
        ctx.upload(g0)
        used, eigs = [], []
        for n in range(steps):
            if dts is None:
                used.append(ctx.step())
            else:
                eigs.append(ctx.evolve_space(n % 2))
                ctx.evolve_time(dts[n])
                used.append(float(dts[n]))
        return ctx.download(), used, eigs
    finally:
        ctx.close()


def run_oracle(meta, g0, steps, dts=None, eigen="closed"):
    cfg = oracle_cfg(meta, eigen)
    g, used = advance(np.copy(g0), cfg, steps, dts=dts)
    return g, used
