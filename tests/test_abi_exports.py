"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol the
header declares, and fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes
import os
import re

import pytest
from conftest import load_pkg, has_gpu, ROOT


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "sdso_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sdso_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    for must in ("sdso_ctx_create", "sdso_make_images", "sdso_tracker_set_ref", "sdso_calc_res_gs", "sdso_track"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    pkg = load_pkg()
    lib = ctypes.CDLL(pkg.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in include/sdso_b200.h but not exported: {missing}"


def test_default_settings_match_reference_preset():
    pkg = load_pkg()
    s = pkg.default_settings()
    # util/settings.cpp:95,102,72,73 and main_dso_pangolin.cpp:325-327 (mode=1)
    assert s.huberTH == 9 and s.coarseCutoffTH == 20 and s.outlierTH == 144 and s.outlierTHSumComponent == 2500
    assert s.affineOptModeA == 0 and s.affineOptModeB == 0
    assert s.trace_GNIterations == 3 and s.minTraceTestRadius == 2


@pytest.mark.skipif(has_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    pkg = load_pkg()
    with pytest.raises(pkg.SdsoError):
        pkg.Context(1232, 368, (718.0, 718.0, 600.0, 180.0))


def test_product_does_not_reference_the_oracle():
    """The product tree must never include, link or import anything under oracle/."""
    prod = os.path.join(ROOT, "stereo-dso-g2o_b200")
    for dp, _, files in os.walk(prod):
        if os.path.basename(dp) == "build":
            continue
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".hpp", ".cpp", ".py")) or f == "Makefile":
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in txt.lower(), f"{os.path.join(dp, f)} mentions the oracle"
