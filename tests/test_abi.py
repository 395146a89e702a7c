"""The C-ABI shared library loads and exports every symbol include/trb.h declares (no compute
calls: there is no GPU here), and the Python binding table matches the header."""
import ctypes
import os
import re

import tinyrenderder_b200 as trb
from tinyrenderder_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "trb.h")).read()
    return sorted(set(re.findall(r"TRB_FN\((\w+)\)\s*\(", text)))


def test_header_and_binding_table_agree():
    assert declared_functions() == sorted(capi.SIGNATURES)


def test_cuda_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(trb.CUDA_LIB)
    for name in declared_functions():
        assert hasattr(lib, "trb_" + name), "libtrb.so does not export trb_" + name
    lib.trb_backend_name.restype = ctypes.c_char_p
    assert lib.trb_backend_name() == b"cuda-sm100a"


def test_oracles_export_the_same_abi(built):
    for rel in ("oracle/libtrb_port.so", "oracle/_ref/libtrb_ref.so"):
        p = os.path.join(ROOT, rel)
        if not os.path.exists(p):
            continue
        lib = ctypes.CDLL(p)
        for name in declared_functions():
            assert hasattr(lib, "orc_" + name), "%s does not export orc_%s" % (rel, name)


def test_cuda_library_is_sm100a_only(built):
    import subprocess
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", trb.CUDA_LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out


def test_no_fallback_without_device(built):
    """without a usable sm_100 device trb_create must fail instead of silently running elsewhere"""
    api = trb.load_cuda()
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    try:
        trb.Renderer(api)
    except trb.TrbError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("trb_create succeeded without a GPU")


def test_product_package_does_not_touch_the_oracle():
    pkg = os.path.join(ROOT, "tinyrenderder_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "libtrb_port" not in text and "libtrb_ref" not in text and "oracle/" not in text.replace(
                    "the CPU oracle", ""), "%s references the oracle" % f
