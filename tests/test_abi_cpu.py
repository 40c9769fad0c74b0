"""CPU tests of the boundary: the C-ABI library loads, exports every declared symbol, and fails loudly without a GPU."""
import ctypes as C
import os
import re
import subprocess
import numpy as np
import pytest
from conftest import ROOT, PKG_DIR


def declared_functions(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b((?:ggd|lps)_[a-z0-9_]+)\s*\(", txt)))


@pytest.mark.parametrize("header", ["ggd_train.h", "lps_b200.h"])
def test_library_exports_every_declared_symbol(pkg, header):
    L = pkg.load_library()
    names = declared_functions(header)
    assert len(names) >= 8
    for n in names:
        assert hasattr(L, n), "libggd_b200.so does not export %s declared in include/%s" % (n, header)


def test_sass_is_blackwell_native():
    """the shipped library contains tcgen05 MMAs, TMEM loads and TMA loads (B200_PROFILING.md mnemonics)"""
    so = os.path.join(PKG_DIR, "libggd_b200.so")
    try:
        sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, timeout=300).stdout
    except FileNotFoundError:
        pytest.skip("cuobjdump not available")
    if not sass:
        pytest.skip("no SASS dump")
    assert "UTCHMMA" in sass and "LDTM" in sass and "UTMALDG" in sass
    assert "HMMA." not in sass.replace("UTCHMMA", "")      # no legacy mma.sync tensor path


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from oracle import oracle as O
    W, b = O.init_weights([8, 6, 5])
    with pytest.raises(pkg.GGDError, match="no CUDA device"):
        pkg.BP_GPU(0, 0, 3, [8, 6, 5], 128, 0.1, 0.9, 0.0, W, b, 1.5, 1)
    with pytest.raises(pkg.LPSError, match="no CUDA device"):
        pkg.Wav2LPS(0)
    assert pkg.lps_nframes(43264) == 168      # pure host arithmetic, no device needed


def test_product_does_not_import_oracle():
    """the product path must never route through oracle/ (or any CPU fallback)"""
    for dirpath, _, files in os.walk(PKG_DIR):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".sh")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower().replace("oracle/_ref", ""), os.path.join(dirpath, f)


def test_config_struct_layout(pkg):
    """ctypes mirror of ggd_config matches the C header field order"""
    from se_ml_b200.bp_gpu import _Config
    txt = open(os.path.join(ROOT, "include", "ggd_train.h")).read()
    body = re.search(r"typedef struct ggd_config \{(.*?)\} ggd_config;", txt, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        decl = re.sub(r"^(const\s+)?(int|float|void)\s*\*?", "", decl)
        for nm in decl.split(","):
            fields.append(re.sub(r"\[.*\]", "", nm).replace("*", "").strip())
    assert fields == [f[0] for f in _Config._fields_]
