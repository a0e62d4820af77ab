"""schemes/ppm.py:111-170 at function level: slope flattener and artificial viscosity.

The flattener is pinned by tests/golden/f_ppm_flattener.npz (outputs of the unmodified reference's
``ppm.apply_flattener`` on evolved states, made by tests/golden/make_function_golden.py): the oracle and the kernel
(host simulation here, the sm_100a library under ``-m gpu``) must reproduce every vector bit for bit.  The artificial
viscosity has no reference output (the reference raises for every grid but an 8-cell 1D one, recorded in
f_ppm_flattener.json); the kernel is compared with the oracle's cell-wise restatement, and 2D input must raise.
"""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

VECTORS = np.load(os.path.join(GOLDEN, "f_ppm_flattener.npz"))
META = json.load(open(os.path.join(GOLDEN, "f_ppm_flattener.json")))
KEYS = sorted(k[:-2] for k in VECTORS.files if k.endswith("|w"))


@pytest.mark.parametrize("key", KEYS)
def test_oracle_flattener_equals_reference(key):
    from oracle.reconstruct import ppm_flattener
    cid, axis, _ = key.split("|")
    w, chi = VECTORS[key + "|w"], VECTORS[key + "|chi"]
    with np.errstate(all="ignore"):
        got = ppm_flattener(np.copy(w), int(axis), META["boundary"][cid])
    assert np.array_equal(got[..., 0], chi, equal_nan=True)
    assert all(np.array_equal(got[..., k], chi, equal_nan=True) for k in range(8))


def _check_flattener(lib, key):
    from astrea_b200 import ppm
    cid, axis, _ = key.split("|")
    w, chi = VECTORS[key + "|w"], VECTORS[key + "|chi"]
    eta = ppm.apply_flattener(w, int(axis), META["boundary"][cid], _lib=lib)
    assert eta.shape == w.shape
    assert all(np.array_equal(eta[..., k], chi, equal_nan=True) for k in range(8))


def _check_viscosity(lib):
    from collections import namedtuple
    from astrea_b200 import ppm
    from oracle.reconstruct import ppm_artificial_viscosity_cellwise
    SV = namedtuple("sv", "boundary gamma dx dimension")
    for key in KEYS:
        cid, axis, tag = key.split("|")
        w = VECTORS[key + "|w"]
        sv = SV(META["boundary"][cid], 1.4, 1.0 / w.shape[0], w.ndim - 1)
        if w.ndim == 3:
            with pytest.raises(ValueError):
                ppm.apply_artificial_viscosity(w, int(axis), sv, _lib=lib)
            continue
        with np.errstate(all="ignore"):
            want = ppm_artificial_viscosity_cellwise(np.copy(w), int(axis), sv)
        got = ppm.apply_artificial_viscosity(w, int(axis), sv, _lib=lib)
        assert np.array_equal(got, want, equal_nan=True), key
        if not tag:
            assert np.abs(want).max() > 0 or "shu" in cid or "rj" in cid      # shocks compress: the coefficient is active
    # custom coefficients reach the kernel
    w = VECTORS["x_sod_ppm_hllc_ssprk54|0||w"]
    sv = SV("edge", 1.4, 1.0 / w.shape[0], 1)
    a = ppm.apply_artificial_viscosity(w, 0, sv, viscosity_determinants=(.6, .3), _lib=lib)
    b = ppm.apply_artificial_viscosity(w, 0, sv, _lib=lib)
    assert np.array_equal(a, 2 * b)
    ppm.release()


@pytest.mark.parametrize("key", KEYS)
def test_hostsim_flattener_equals_reference(hostsim_lib, key):
    _check_flattener(hostsim_lib, key)


def test_hostsim_flattener_custom_determinants(hostsim_lib):
    from astrea_b200 import ppm
    from oracle.reconstruct import ppm_flattener
    key = "c2_ll3_ppm_hllc_ssprk33|1|"
    w = VECTORS[key + "|w"]
    knobs = (.2, .6, .9)
    with np.errstate(all="ignore"):
        want = ppm_flattener(np.copy(w), 1, "wrap", knobs)
    assert np.array_equal(ppm.apply_flattener(w, 1, "wrap", knobs, _lib=hostsim_lib), want, equal_nan=True)
    ppm.release()


def test_hostsim_viscosity(hostsim_lib):
    _check_viscosity(hostsim_lib)


def test_reference_viscosity_raises_is_recorded():
    """Why there is no golden vector for apply_artificial_viscosity."""
    assert all(v and "broadcast" in v for v in META["viscosity_raises"].values())


@pytest.mark.gpu
@pytest.mark.parametrize("key", KEYS)
def test_gpu_flattener_equals_reference(key):
    from astrea_b200 import _native
    _check_flattener(_native.device_library(), key)


@pytest.mark.gpu
def test_gpu_viscosity():
    from astrea_b200 import _native
    _check_viscosity(_native.device_library())
