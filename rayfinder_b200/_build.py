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
CXX_FLAGS = ["-std=c++20", "-O2", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-Wall", "-Wextra",
             f'-DRF_DATA_DIR="{DATA}"']


def _newer(target: Path, sources: list[Path]) -> bool:
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(s.stat().st_mtime <= t for s in sources)


def build(force: bool = False, verbose: bool = False) -> Path:
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
            cmd = [NVCC, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
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
    cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *objs]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
