"""Build libinstarevive_b200.so in-tree with nvcc for sm_100a (no torch extension machinery: the library is plain C ABI).

Usage: python -m instarevive_b200.csrc.build [--force] [--verbose]
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
SOURCES = ["api.cu", "gemm.cu", "attention.cu", "attention_tc.cu", "xattention_tc.cu", "elementwise.cu", "dit.cu", "vae.cu", "tiles.cu", "swinir.cu"]
LIB = HERE / "libinstarevive_b200.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]
if os.environ.get("IR_DEBUG") == "1":   # experiment build: the library honours the IR_* A/B environment switches
    FLAGS.append("-DIR_DEBUG")


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    srcs = [HERE / s for s in SOURCES if (HERE / s).exists()]
    deps = srcs + sorted(HERE.glob("*.cuh")) + [HERE.parent.parent / "include" / "instarevive_b200.h"]
    stamp = HERE / "build" / "digest.txt"
    digest = _digest(deps)
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == digest:
        return LIB
    (HERE / "build").mkdir(exist_ok=True)

    def compile_one(src: Path):
        obj = HERE / "build" / (src.stem + ".o")
        cmd = [NVCC, *FLAGS, "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        (HERE / "build" / (src.stem + ".ptxas.log")).write_text(r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [NVCC, "-shared", "-o", str(LIB), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(digest)
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
