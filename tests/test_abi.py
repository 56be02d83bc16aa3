"""The C-ABI library loads on a machine without a GPU and exports exactly what
include/ab200.h declares; with no device every compute path fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from artemis_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ab200.h")


def _declared():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ab200_[a-z0-9_]+)\s*\(", txt)))


@pytest.mark.parametrize("variant", ["fast", "strict"])
def test_library_exports_every_declared_symbol(variant):
    L = capi.load(variant)
    names = _declared()
    assert len(names) >= 39
    for n in names:
        assert hasattr(L, n), f"{n} declared in ab200.h but not exported by the {variant} library"
    assert sorted(capi.SYMBOLS) == names, "capi.SYMBOLS is out of sync with include/ab200.h"
    assert L.ab200_abi_version() == 12


def test_header_is_plain_c():
    """The boundary is extern "C" with plain pointers and sizes: it must compile as C99."""
    import subprocess
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "t.c")
        with open(src, "w") as fh:
            fh.write('#include "ab200.h"\nint main(void){ab200_grid_desc g; (void)g; return 0;}\n')
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I",
                               os.path.join(ROOT, "include"), "-c", src, "-o",
                               os.path.join(td, "t.o")])


def test_every_entry_point_cites_the_reference():
    txt = open(HEADER).read()
    for cite in ("src/gas/gas.cpp:473-494", "fluid_fluxes.hpp:76-213",
                 "artemis_integrator.hpp:56-110", "fluid_fluxes.hpp:298-420",
                 "fill_derived.cpp:29-75", "fill_derived.cpp:81-167", "fill_derived.cpp:172-277",
                 "src/gas/gas.cpp:391-468", "boundary_communication.cpp:95-140",
                 "P:driver/driver.cpp:210-269"):
        assert cite in txt, cite


def test_no_cpu_fallback_without_a_device():
    L = capi.load("fast")
    if L.ab200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    ctx = C.c_void_p()
    rc = L.ab200_create(C.byref(ctx), 0, None)
    assert rc == 2  # AB200_ECUDA
    assert b"no CPU fallback" in L.ab200_last_error()
    with pytest.raises(capi.AB200Error):
        capi.check(L, rc, "ab200_create")
    # null-context calls are rejected, not crashed
    assert L.ab200_fused_stage(None, 0.0, 1.0, 1.0, 0.1, 0, 1, 0) != 0
    assert L.ab200_calculate_fluxes(None, 0, 0) != 0


def test_product_package_does_not_import_the_oracle():
    """oracle/ is test infrastructure: nothing under artemis_b200/ may reference it."""
    pkg = os.path.join(ROOT, "artemis_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("# oracle", ""), os.path.join(dp, f)
