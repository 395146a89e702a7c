"""tinyrenderder_b200 - B200-native rasterization backend behind the our_gl.h API of
AnnaUshnova/tinyrenderder.  The product is the CUDA library `libtrb.so` (C ABI in
include/trb.h, built from csrc/ by `__graft_entry__.build()`); this package is the thin host
binding used by the tests and bench.py.  There is no CPU fallback."""
from .capi import (Api, Renderer, TrbError, PhongUniforms, ShadowUniforms, Stats, load_cuda, CUDA_LIB,
                   SHADER_FLAT_BARY, SHADER_PHONG, SHADER_EYE, SHADER_DEPTH, SHADER_SHADOW_PHONG,
                   SHADER_GOURAUD, VIS_NONE, VIS_SHADED, comm_init, composite_group)

__all__ = ["Api", "Renderer", "TrbError", "PhongUniforms", "ShadowUniforms", "Stats", "load_cuda", "CUDA_LIB",
           "SHADER_FLAT_BARY", "SHADER_PHONG", "SHADER_EYE", "SHADER_DEPTH", "SHADER_SHADOW_PHONG",
           "SHADER_GOURAUD", "VIS_NONE", "VIS_SHADED", "comm_init", "composite_group"]
