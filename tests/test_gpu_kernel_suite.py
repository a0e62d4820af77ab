"""The whole kernel-logic suite of tests/test_hostsim_parity.py once more, on the sm_100a library on a B200.

tests/test_hostsim_parity.py checks every index map, halo rule, limiter, integrator, download path and call-order
rule against the oracle on the g++ host simulation of the kernel sources (the build container has no GPU).  The
host simulation is test infrastructure; what ships is the CUDA library, so every one of those comparisons is repeated
here through the C ABI of ``astrea_b200/lib/libastrea_b200.so``: the test functions are imported unchanged and the
``hostsim_lib`` fixture they ask for is overridden, for this module, by the device library.  Covers on the GPU
what used to be host-simulation only: the five r-based slope limiters (limiters.py:23-49), the primitive download
and the snapshot layout (astrea.py:47), the face-field download (evolvers.py:73-76), 4- and 6-cell grids, the
step-program contract and the split register update.
"""
import pytest

import test_hostsim_parity as _suite

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hostsim_lib():
    """Overrides the session fixture of conftest.py for the tests collected in this module: the device library."""
    from astrea_b200 import _native
    lib = _native.device_library()       # raises if the CUDA library is missing: there is no fallback
    assert lib.astrea_is_device_build() == 1
    return lib


for _name in dir(_suite):
    if _name.startswith("test_"):
        globals()[_name] = getattr(_suite, _name)
del _name
