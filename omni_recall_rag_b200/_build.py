"""In-tree build of liborr.so (hand-written sm_100a CUDA + the C ABI) with nvcc.

The library is built next to this file so it travels to the GPU box with the repo snapshot;
nothing is JIT-compiled at run time.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
BUILD_DIR = os.path.join(PKG_DIR, "build")
LIB_PATH = os.path.join(PKG_DIR, "liborr.so")

SOURCES = ["orr_api.cu", "orr_scan.cu", "orr_rescore.cu", "orr_exact.cu", "orr_vocab.cu", "orr_synth.cu", "orr_batch.cu", "orr_xchg.cu", "orr_cluster.cu", "orr_textmatch.cu", "orr_text.cpp"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
    "--expt-relaxed-constexpr",
    # fp64/fp32 arithmetic that must match the reference is written with explicit _rn
    # intrinsics; fast-math stays off so exp/sqrt/div keep their IEEE/<=1ulp versions.
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; liborr.so cannot be built (there is no CPU fallback)")


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _fingerprint() -> str:
    h = hashlib.sha256()
    names = sorted(os.listdir(CSRC)) + ["../../include/orr.h"]
    for name in names:
        path = os.path.join(CSRC, name)
        if os.path.isfile(path):
            h.update(name.encode())
            h.update(open(path, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_native(force: bool = False, verbose: bool = False) -> str:
    """Compiles csrc/*.cu for sm_100a and links liborr.so.  Returns the library path."""
    os.makedirs(BUILD_DIR, exist_ok=True)
    stamp = os.path.join(BUILD_DIR, "fingerprint")
    fp = _fingerprint()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp) and open(stamp).read() == fp:
        return LIB_PATH
    nvcc = _nvcc()
    srcs = _sources()

    def compile_one(src: str) -> str:
        obj = os.path.join(BUILD_DIR, os.path.splitext(src)[0] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if src.endswith(".cu") and verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose and r.stderr:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    # the SHARED CUDA runtime: linking cudart_static would embed the runtime's whole entry-point table in liborr.so (symbol names
    # of calls this library never makes)
    cmd = [nvcc, "-shared", "-o", LIB_PATH, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "shared", "-lpthread", "-ldl", "-lrt"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(fp)
    return LIB_PATH


if __name__ == "__main__":
    import sys

    print(build_native(force="--force" in sys.argv, verbose="-v" in sys.argv))
