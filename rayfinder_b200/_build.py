"""Build recipe for librayfinder_b200.so (in-tree, sm_100a only).

    python -m rayfinder_b200._build          # or __graft_entry__.build()

Host files are compiled with g++ -ffp-contract=off, the CUDA file with nvcc -fmad=false: the parity
contract of the path is strict IEEE fp32 with no FMA contraction (DESIGN.md "Arithmetic").
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
DATA = PKG / "data"
LIB = PKG / "librayfinder_b200.so"
OBJ = PKG / "build"

NVCC = os.environ.get("NVCC", shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc")
CXX = os.environ.get("CXX", "g++")

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math", "-Xptxas", "-v",
]
# bvh_build.cu: its single-launch builder passes data between blocks through global memory inside one kernel, so plain global
# loads must not be served from a (non-coherent) L1 line: cache them in L2 only.
EXTRA_NVCC_FLAGS = {"bvh_build.cu": ["-Xptxas", "-dlcm=cg"]}
CXX_FLAGS = ["-std=c++20", "-O2", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-Wall", "-Wextra",
             f'-DRF_DATA_DIR="{DATA}"']


def _newer(target: Path, sources: list[Path]) -> bool:
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(s.stat().st_mtime <= t for s in sources)


def build(force: bool = False, verbose: bool = False, timeline: bool = False) -> Path:
    if timeline:
        return _build_timeline(verbose)
    headers = sorted(CSRC.glob("*.h")) + sorted(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "rayfinder_b200.h"]
    data = sorted(DATA.glob("*.bin"))
    sources = sorted(CSRC.glob("*.cpp")) + sorted(CSRC.glob("*.cu"))
    if not force and _newer(LIB, sources + headers + data + [Path(__file__)]):
        return LIB
    OBJ.mkdir(exist_ok=True)
    objs = []
    for src in sources:
        obj = OBJ / (src.name + ".o")
        if src.suffix == ".cu":
            cmd = [NVCC, *NVCC_FLAGS, *EXTRA_NVCC_FLAGS.get(src.name, []), "-c", str(src), "-o", str(obj)]
        else:
            cmd = [CXX, *CXX_FLAGS, "-c", str(src), "-o", str(obj)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or res.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError(f"compilation of {src.name} failed")
        if src.suffix == ".cu":
            (OBJ / (src.name + ".ptxas.txt")).write_text(res.stderr)
        objs.append(str(obj))
    cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *objs, "-lz"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("link failed")
    return LIB


def _build_timeline(verbose: bool, defines=(), suffix: str = "") -> Path:
    """Instrumented debug build (per-warp timeline of the traversal launches, -DRF_TRACE_TIMELINE) next to the
    product library: librayfinder_b200_timeline.so, selected with RAYFINDER_B200_LIB (tools/trace_timeline.py)."""
    out = PKG / f"librayfinder_b200_timeline{suffix}.so"
    OBJ.mkdir(exist_ok=True)
    obj = OBJ / f"device.timeline{suffix}.o"
    cmd = [NVCC, *NVCC_FLAGS, "-DRF_TRACE_TIMELINE", *[f"-D{d}" for d in defines], "-c", str(CSRC / "device.cu"), "-o", str(obj)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("timeline build failed")
    objs = [str(obj)] + [str(OBJ / (src.name + ".o")) for src in sorted(CSRC.glob("*.cpp")) + sorted(CSRC.glob("*.cu")) if src.name != "device.cu"]
    res = subprocess.run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(out), *objs, "-lz"], capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("link failed")
    return out


if __name__ == "__main__":
    if "--timeline" in sys.argv:
        build()
        defines = [a[2:] for a in sys.argv if a.startswith("-D")]
        suffix = next((a.split("=", 1)[1] for a in sys.argv if a.startswith("--suffix=")), "")
        print(_build_timeline(True, defines, suffix))
        sys.exit(0)
    print(build(force="--force" in sys.argv, verbose=True))
