"""C-ABI checks that need no GPU: the library loads, exports every symbol include/fhe_sign_cuda.h
declares, and fails loudly (no CPU fallback) when no device is present."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(hdr):
    names = re.findall(r"^(?:fsc_status|const char \*)\s*(fsc_[a-z0-9_]+)\s*\(", hdr, flags=re.M)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    import fhe_sign_b200 as fsb
    L = fsb.load_library()
    hdr = open(os.path.join(ROOT, "include", "fhe_sign_cuda.h")).read()
    declared = _declared(hdr)
    assert len(declared) >= 20
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing


def test_binding_lists_match_header():
    from fhe_sign_b200.capi import EXPORTS
    hdr = open(os.path.join(ROOT, "include", "fhe_sign_cuda.h")).read()
    declared = set(_declared(hdr))
    assert declared == set(EXPORTS)


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import fhe_sign_b200 as fsb
    with pytest.raises(fsb.FscError) as ei:
        fsb.Context(fsb.Params.preset("toy"))
    assert ei.value.code == 4      # FSC_ERR_CUDA


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "fhe_sign_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(d, f)).read()
                assert "liborc" not in txt and "tfhe_oracle" not in txt and "from oracle" not in txt, f
