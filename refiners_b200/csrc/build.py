"""Build librefiners_b200.so in-tree with nvcc for sm_100a (no torch, no CPU fallback objects).

Usage: python -m refiners_b200.csrc.build [--force] [--verbose]
Objects are cached under csrc/build/ keyed on source+header mtimes, so rebuilding after
touching one kernel file recompiles only that file.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
SOURCES = ["api.cu", "simt_gemm.cu", "simt_attention.cu", "norm_kernels.cu", "tc_gemm.cu", "tc_attention.cu", "tc_attention2.cu", "tc_attention_short.cu", "tc_attention_win.cu", "attn_probs.cu", "sam_attention.cu"]
HEADERS = [HERE / "common.cuh", HERE / "tc_ptx.cuh", ROOT / "include" / "refiners_b200.h"]
LIB = HERE / "librefiners_b200.so"
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v" if os.environ.get("RB200_PTXAS_V") else "-O3",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(exe).exists():
        raise RuntimeError("nvcc not found: refiners_b200 needs the CUDA toolkit to build its kernels")
    return exe


def _stale(obj: Path, src: Path) -> bool:
    if not obj.exists():
        return True
    t = obj.stat().st_mtime
    return any(dep.stat().st_mtime > t for dep in [src, *HEADERS])


def build(force: bool = False, verbose: bool = False) -> Path:
    out = HERE / "build"
    out.mkdir(exist_ok=True)
    jobs = []
    for name in SOURCES:
        src, obj = HERE / name, out / (name + ".o")
        if force or _stale(obj, src):
            jobs.append((src, obj))

    def compile_one(job: tuple[Path, Path]) -> None:
        src, obj = job
        cmd = [nvcc(), *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        if verbose:
            print(" ".join(cmd), flush=True)
        res = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or res.returncode:
            sys.stderr.write(res.stdout + res.stderr)
        if res.returncode:
            raise RuntimeError(f"nvcc failed on {src.name}")

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as pool:
        list(pool.map(compile_one, jobs))

    objs = [out / (n + ".o") for n in SOURCES]
    if force or jobs or not LIB.exists():
        cmd = [nvcc(), "-shared", "-cudart", "shared", "-o", str(LIB), *map(str, objs),
               "-Xlinker", "-rpath,/usr/local/cuda/lib64"]
        if verbose:
            print(" ".join(cmd), flush=True)
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
