"""Build libditb200.so (hand-written sm_100a CUDA, C-ABI) in-tree with nvcc.

    python -m fast_dit_b200.build [--force] [--verbose]

The library has no torch dependency: it is plain CUDA runtime code behind
include/ditb200.h.  nvcc cross-compiles sm_100a without a GPU, so this runs in
the CPU-only build container; the resulting .so travels to the GPU box.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OUT_DIR = PKG / "lib"
OBJ_DIR = PKG / "lib" / "obj"
LIB = OUT_DIR / "libditb200.so"
STAMP = OUT_DIR / "libditb200.stamp"

SOURCES = [
    "api.cu",
    "elementwise.cu",
    "diffusion.cu",
    "gemm_simt.cu",
    "gemm_tc.cu",
    "attention.cu",
    "attention_tc.cu",
    "backward.cu",
    "optim.cu",
]

NVCC_FLAGS = [
    "-O3",
    "-std=c++17",
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler",
    "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas",
    "-v",
]
# measurement variants (tools/ab.sh): e.g. DITB200_NVCC_EXTRA=-DDITB200_NO_PDL builds the library without programmatic dependent launch
# variant of the hot kernels; part of the fingerprint, so switching it rebuilds
NVCC_FLAGS += os.environ.get("DITB200_NVCC_EXTRA", "").split()


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libditb200 cannot be built")
    return nvcc


def _fingerprint() -> str:
    h = hashlib.sha256()
    files = sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "ditb200.h"]
    for f in files:
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile(src: str, verbose: bool) -> str:
    obj = OBJ_DIR / (Path(src).stem + ".o")
    cmd = [_nvcc(), *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    (OBJ_DIR / (Path(src).stem + ".log")).write_text(log)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{log}")
    if verbose:
        print(log)
    return str(obj)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every kernel for sm_100a and link libditb200.so; no-op if up to date."""
    fp = _fingerprint()
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == fp:
        return LIB
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    srcs = [s for s in SOURCES if (CSRC / s).exists()]
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), srcs))
    cmd = [_nvcc(), "-shared", "-o", str(LIB), *objs, "-gencode", "arch=compute_100a,code=sm_100a",
           "-Xcompiler", "-fPIC", "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}{r.stderr}")
    STAMP.write_text(fp)
    return LIB


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    lib = build(force=a.force, verbose=a.verbose)
    print(lib)


if __name__ == "__main__":
    sys.exit(main())
