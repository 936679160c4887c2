"""Compile libjmmgpu.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libjmmgpu.so"


def _stale() -> bool:
    if not LIB.exists() or not (PKG / "bin" / "jmm_run").exists():
        return True
    t = LIB.stat().st_mtime
    srcs = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("host/*")) + \
        [PKG.parent / "include" / "jmm_gpu.h", CSRC / "Makefile"]
    return any(s.exists() and s.stat().st_mtime > t for s in srcs)


def build(force: bool = False, verbose: bool = False) -> Path:
    if force or _stale():
        env = dict(os.environ)
        env.setdefault("NVCC", "/usr/local/cuda/bin/nvcc" if Path("/usr/local/cuda/bin/nvcc").exists() else "nvcc")
        args = ["make", "-j", "4", "-C", str(CSRC), "all"] + (["-B"] if force else [])
        r = subprocess.run(args, env=env, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            print(r.stdout)
            print(r.stderr)
        if r.returncode != 0:
            raise RuntimeError("building libjmmgpu.so failed (see output above)")
    return LIB
