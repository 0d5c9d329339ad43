"""No-GPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/ti_b200.h declares,
reports its ABI version, and every compute entry point fails LOUDLY without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ti_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ti_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import turboinfer_b200 as tb
    lib = tb.lib()
    names = declared_symbols()
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.ti_b200_abi_version() == 1


def test_no_cpu_fallback_without_device():
    import turboinfer_b200 as tb
    lib = tb.lib()
    n = C.c_int(-1)
    lib.ti_b200_device_count(C.byref(n))
    if n.value > 0:
        return  # on a GPU box this check is the C++ test's job (CUDA_VISIBLE_DEVICES="")
    assert lib.ti_b200_init(0) != 0
    assert b"no CPU fallback" in lib.ti_b200_last_error()
    x = np.ones(8, np.float32)
    y = np.zeros(8, np.float32)
    fp = C.POINTER(C.c_float)
    lib.ti_b200_silu.argtypes = [fp, fp, C.c_size_t]
    assert lib.ti_b200_silu(x.ctypes.data_as(fp), y.ctypes.data_as(fp), 8) != 0
    assert b"not initialised" in lib.ti_b200_last_error()
    assert not y.any()


def test_product_never_imports_the_oracle():
    for base, _, files in os.walk(os.path.join(ROOT, "turboinfer_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                src = open(os.path.join(base, f)).read()
                assert "import oracle" not in src and "oracle/" not in src and "libti_oracle" not in src and "libti_ref" not in src, f
