"""The C-ABI library loads and exports every function include/lbmpc.h declares; without a GPU it refuses
to work instead of falling back to the CPU.  No compute calls here."""
import ctypes as C
import os
import re

import pytest

from lbmpc_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    txt = open(os.path.join(ROOT, "include", "lbmpc.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(lbmpc_[a-z_]+)\s*\(", txt)))


def test_header_declares_the_boundary():
    names = declared_functions()
    for f in ("lbmpc_create", "lbmpc_solve_batch", "lbmpc_oracle_apply", "lbmpc_closed_loop", "lbmpc_destroy",
              "lbmpc_last_error"):
        assert f in names


def test_library_exports_every_declared_symbol():
    lib = capi.load_library()
    for f in declared_functions():
        assert hasattr(lib, f), f"{f} declared in include/lbmpc.h but not exported"
    assert lib.lbmpc_version().decode().startswith("lbmpc_b200")


def test_struct_layouts_match_header():
    """ctypes mirrors of lbmpc_model / lbmpc_config list the header's fields in the header's order."""
    txt = open(os.path.join(ROOT, "include", "lbmpc.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    for struct, cls in (("lbmpc_model", capi.LbmpcModel), ("lbmpc_config", capi.LbmpcConfig)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct, struct), txt, re.S).group(1)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            names = decl.replace("const", "").replace("double", "").replace("int32_t", "").replace("int64_t", "")
            fields += [n.strip().lstrip("*") for n in names.split(",")]
        assert fields == [f[0] for f in cls._fields_], struct


def test_no_cpu_fallback_without_gpu(models):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.LbmpcError) as e:
        capi.Solver(models["LBMPC"], "C", "LBMPC", 50)
    assert "CUDA" in str(e.value)
