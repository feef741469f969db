"""CPU tests of the drop-in boundary: libb2rns.so loads, exports every symbol include/b2rns.h declares, validates
arguments on the host, and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "b2rns.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(b2):
    L = b2.lib()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), "libb2rns.so does not export %s" % n
    out = subprocess.run(["nm", "-D", "--defined-only", b2.LIB_PATH], capture_output=True, text=True).stdout
    for n in names:
        assert re.search(r"\bT %s\b" % n, out), "%s is not an extern \"C\" text symbol" % n
    # python binding covers the whole header
    assert set(names) <= set(b2.exported_symbols()), sorted(set(names) - set(b2.exported_symbols()))


def test_signatures_are_plain_c(b2):
    src = open(os.path.join(ROOT, "include", "b2rns.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)      # comments may cite the reference's C++ exceptions
    assert "torch" not in src and "std::" not in src and "#include <cuda" not in src


def test_no_product_dependency_on_oracle():
    """the shipped library must not link, load or mention anything under oracle/"""
    out = subprocess.run(["ldd", os.path.join(ROOT, "troy-nova_b200", "libb2rns.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out and "troy_ref" not in out and "ref_shim" not in out
    for dirpath, _, files in os.walk(os.path.join(ROOT, "troy-nova_b200")):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".cpp", ".py")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in txt and "rns_oracle" not in txt and "ref_shim" not in txt, os.path.join(dirpath, f)


def test_host_side_prime_generation_matches_oracle(b2, po):
    for n, bits in ((8192, [40, 40, 40]), (16384, [60, 40, 40, 60]), (65536, [60] + [40] * 19 + [60]), (32768, [60] + [40] * 7 + [60])):
        assert b2.coeff_modulus_create(n, bits) == [int(v) for v in po.coeff_modulus_create(n, bits)]
    with pytest.raises(b2.B2Error):
        b2.coeff_modulus_create(8192, [70])


def test_fails_loudly_without_gpu(b2):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert b2.lib().b2_device_count() == 0
    with pytest.raises(b2.B2Error) as ei:
        b2.Context(12, b2.coeff_modulus_create(4096, [40, 40]), 0, b2.CKKS)
    assert ei.value.status == 3 and "no CPU fallback" in str(ei.value)     # B2_ERR_NO_DEVICE
