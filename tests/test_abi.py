"""CPU-only: the C-ABI library loads, exports every symbol include/stereo_b200.h declares,
reports the reference's macro values as defaults, and fails loudly without a GPU."""
import ctypes as C
import os
import re

import pytest

import stereo_matching_cuda_b200 as S
from stereo_matching_cuda_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "stereo_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sb200_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported():
    lib = C.CDLL(S.lib_path())
    names = header_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/stereo_b200.h but not exported"


def test_binding_covers_header():
    lib = S.load_library()
    bound = set(lib._sb_signatures)
    assert set(header_symbols()) <= bound


def test_default_params_are_the_reference_macros():
    p = api.default_params()  # SystemIncludes.h:6-24
    assert (p.dmin, p.dmax, p.radius, p.d_lr) == (-15, 0, 9, 0)
    assert p.eps == 6.5025 and p.r_w == 0.299 and p.g_w == 0.587 and p.b_w == 0.0721
    assert abs(p.alpha - 0.9) < 1e-7 and p.th_color == 7.0 and p.th_grad == 2.0
    assert p.size_d == 16 and p.guide_mode == S.GUIDE_GRAY


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(S.StereoB200Error, match="no CUDA device"):
        S.Context(0)


def test_product_never_imports_oracle():
    """The product path must not load, link or include anything under oracle/."""
    pkg = os.path.join(ROOT, "stereo_matching_cuda_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                for bad in ("liboracle", "_oracle", "libref", "oracle/", "stereo_oracle", "import oracle"):
                    if bad == "oracle/" and "see oracle/stereo_oracle.c" in src:
                        src = src.replace("see oracle/stereo_oracle.c", "")
                    assert bad not in src, f"{f} mentions {bad}"
