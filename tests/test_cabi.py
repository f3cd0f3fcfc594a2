"""The C-ABI library loads on a machine without a GPU, exports every declared symbol, and its
compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_every_declared_symbol(evr):
    hdr = "".join(open(os.path.join(ROOT, "include", f)).read() for f in sorted(os.listdir(os.path.join(ROOT, "include"))))
    declared = sorted(set(re.findall(r"\b(evr_sg4_[a-zA-Z_0-9]+)\s*\(", hdr)))
    assert declared, "no declarations found"
    L = C.CDLL(evr.lib.SO_PATH)
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/*.h but not exported"
    assert sorted(evr.lib.EXPORTS) == declared


def test_version_and_error_string(evr):
    L = evr.lib.lib()
    assert L.evr_sg4_version() >= 100
    b, e = C.c_int(), C.c_int()
    assert L.evr_sg4_ini_iGs(10, 0, 0, C.byref(b), C.byref(e)) != 0
    assert b"ini_iGs" in L.evr_sg4_last_error()


def test_tables_builder_rejects_bad_input(evr):
    L = evr.lib.lib()
    h = C.c_void_p()
    nq = np.array([[1, 3]], dtype=np.int32)
    assert L.evr_sg4_tables_build(C.byref(h), 0, 1, 1, nq.ctypes.data, nq.ctypes.data) != 0
    bad = np.array([[3, 1]], dtype=np.int32)          # nb decreasing with L
    assert L.evr_sg4_tables_build(C.byref(h), 1, 1, 1, nq.ctypes.data, bad.ctypes.data) != 0


def test_no_cpu_fallback(evr):
    """Without a CUDA device plan creation must fail with a message, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    basis, op = evr.workloads.henon_heiles(3, 2)
    with pytest.raises(evr.EvrSg4Error, match="CUDA|cuda"):
        op.apply_host(np.zeros(basis.nb))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "elvibrot-tnumtana_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".f90")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in txt.lower().replace("no oracle", ""), f"{f} mentions the oracle"


def test_python_constants_match_the_headers(evr):
    """Enumerators and macros that elvibrot-tnumtana_b200/lib.py restates must equal include/*.h."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    h = open(os.path.join(root, "include", "evr_sg4.h")).read()
    enums = {m.group(1): int(m.group(2)) for m in re.finditer(r"\b(EVR_(?:INFO|TAB)_[A-Z0-9_]+)\s*=\s*(\d+)", h)}
    assert enums, "no enumerators found in include/evr_sg4.h"
    for name, value in enums.items():
        py = name[len("EVR_"):]
        assert hasattr(evr.lib, py), f"lib.py lacks {py}"
        assert getattr(evr.lib, py) == value, (name, value, getattr(evr.lib, py))
    c = open(os.path.join(root, "include", "evr_sg4_comm.h")).read()
    peers = int(re.search(r"#define\s+EVR_SG4_MAX_PEERS\s+(\d+)", c).group(1))
    assert re.search(r"#define\s+EVR_SG4_FLAG_WORDS\s+\(2 \* EVR_SG4_MAX_PEERS \+ 2\)", c)
    assert evr.lib.FLAG_WORDS == 2 * peers + 2
