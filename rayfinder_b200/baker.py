"""GLB / glTF -> .pt scene baker: ``PtFormat(gltfPath)`` and ``pt-format-tool`` of the reference
(pt-format/pt_format.cpp:20-151, pt-format-tool/main.cpp:14-35) — the step before the render path.

The work is done by the C++ baker behind the C-ABI (``rf_bake_gltf``, rayfinder_b200/csrc/host_baker.cpp + host_image.cpp):
glTF parsing, node transforms in glm's fp32 operation order, stb_image's integer PNG / baseline-JPEG decoding, mesh sort,
triangle flattening, ``buildBvh`` and the arrays of the container.  This module is its Python face.

    python -m rayfinder_b200.baker <input_gltf_file> [output.pt]
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import capi
from .api import PtFormat
from .capi import check, lib


def bake(gltf_path) -> PtFormat:
    """``PtFormat(gltfPath)``."""
    handle = C.c_void_p()
    check(lib().rf_bake_gltf(str(gltf_path).encode(), C.byref(handle)))
    return PtFormat._from_handle(handle)


def texture_from_memory(data: bytes) -> np.ndarray:
    """``Texture::fromMemory`` (common/texture.cpp:12-52): (height, width) uint32 BGRA8 of an encoded PNG / baseline JPEG."""
    buf = (C.c_char * len(data)).from_buffer_copy(data) if data else (C.c_char * 1)()
    w, h = C.c_uint32(), C.c_uint32()
    check(lib().rf_texture_from_memory(C.cast(buf, C.c_void_p), len(data), None, 0, C.byref(w), C.byref(h)))
    out = np.zeros((h.value, w.value), dtype="<u4")
    check(lib().rf_texture_from_memory(C.cast(buf, C.c_void_p), len(data), C.c_void_p(out.ctypes.data), out.size, C.byref(w), C.byref(h)))
    return out


def main(argv=None) -> int:  # pt-format-tool/main.cpp:14-35
    import sys

    argv = sys.argv[1:] if argv is None else argv
    if len(argv) not in (1, 2):
        print("Usage: python -m rayfinder_b200.baker <input_gltf_file> [output.pt]")
        return 0
    src = Path(argv[0])
    dst = Path(argv[1]) if len(argv) == 2 else src.with_suffix(".pt")
    try:
        bake(src).save(dst)
    except capi.RayfinderError as e:  # the reference prints the exception and returns 1 (main.cpp:27-33)
        print(e, file=sys.stderr)
        return 1
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
