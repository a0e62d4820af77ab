"""The C-ABI shared library: builds for sm_100a, loads without a GPU, exports what include/astrea_b200.h declares."""
import ctypes
import os
import re

import pytest

from conftest import ROOT, _has_gpu


def _declared():
    text = open(os.path.join(ROOT, "include", "astrea_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(astrea_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree():
    from astrea_b200 import _native
    assert _declared() == sorted(_native.EXPORTS)


def test_device_library_exports_every_symbol(device_lib_path):
    lib = ctypes.CDLL(device_lib_path)
    for name in _declared():
        assert hasattr(lib, name), name
    lib.astrea_is_device_build.restype = ctypes.c_int
    assert lib.astrea_is_device_build() == 1


def test_device_library_contains_sm100a_code(device_lib_path):
    import subprocess
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", device_lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_hostsim_library_is_not_a_device_build(hostsim_lib):
    assert hostsim_lib.astrea_is_device_build() == 0


@pytest.mark.skipif(_has_gpu(), reason="checks the failure mode on a machine without a GPU")
def test_no_cpu_fallback_without_gpu(device_lib_path):
    """Without a CUDA device the product library refuses to create a context (there is no CPU path)."""
    from astrea_b200 import _native as N
    from astrea_b200.selectors import make_cfg
    cfg = make_cfg(dimension=1, cells=64, boundary="edge", gamma=1.4, dx=1 / 64, cfl=.5, subgrid="plm", solver="lf",
                   timestep="ssprk(2,2)")
    with pytest.raises(N.AstreaError) as err:
        N.Context(cfg)
    assert "CUDA" in str(err.value)


def test_package_refuses_non_device_library(monkeypatch, hostsim_lib):
    from astrea_b200 import _native as N, build
    monkeypatch.setattr(N, "DEVICE_LIB", build.HOSTSIM_LIB)
    monkeypatch.setattr(N, "_device_lib", None)
    with pytest.raises(ImportError):
        N.device_library()


def test_struct_layouts_match_the_header(tmp_path):
    """The ctypes mirrors of astrea_cfg / astrea_region / astrea_init_spec have the size and field offsets a C compiler
    gives the declarations of include/astrea_b200.h (the header must also compile as plain C)."""
    import subprocess
    from astrea_b200 import _native as N
    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "astrea_b200.h"\n'
        'int main(void) {\n'
        '  printf("%zu %zu %zu %zu\\n", sizeof(astrea_cfg), offsetof(astrea_cfg, gamma), offsetof(astrea_cfg, nx_global), offsetof(astrea_cfg, flags));\n'
        '  printf("%zu %zu %zu\\n", sizeof(astrea_region), offsetof(astrea_region, a), offsetof(astrea_region, state));\n'
        '  printf("%zu %zu %zu %zu\\n", sizeof(astrea_init_spec), offsetof(astrea_init_spec, background), offsetof(astrea_init_spec, nregions), offsetof(astrea_init_spec, regions));\n'
        '  printf("%zu %zu %zu\\n", sizeof(astrea_init_profile), offsetof(astrea_init_profile, along), offsetof(astrea_init_profile, values));\n'
        '  return 0;\n}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    got = [int(x) for x in out]
    want = [ctypes.sizeof(N.Cfg), N.Cfg.gamma.offset, N.Cfg.nx_global.offset, N.Cfg.flags.offset,
            ctypes.sizeof(N.Region), N.Region.a.offset, N.Region.state.offset,
            ctypes.sizeof(N.InitSpec), N.InitSpec.background.offset, N.InitSpec.nregions.offset, N.InitSpec.regions.offset,
            ctypes.sizeof(N.InitProfile), N.InitProfile.along.offset, N.InitProfile.values.offset]
    assert got == want
