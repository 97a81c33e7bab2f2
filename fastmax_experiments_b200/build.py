"""In-tree build of libfastmax_b200.so (replaces the reference's setup_fast_cuda.py:24-34,
whose sources were never vendored).  nvcc cross-compiles sm_100a without a GPU."""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfastmax_b200.so")
# developer library: the product objects plus csrc/dev/ (micro-benchmarks and the known-answer test of the
# sm_100a primitives; tools/*.py and tests/test_gpu_umma.py load it) -- kept out of the product .so
DEV_LIB_NAME = "libfastmax_b200_dev.so"
OBJ = os.path.join(HERE, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--use_fast_math", "-Xcompiler", "-fPIC", "-Xcompiler", "-O3",
]
# developer switch, e.g. FASTMAX_NVCC_EXTRA="-DFM_TIMING" (in-kernel cycle counters, tools/timing.py)
NVCC_FLAGS += os.environ.get("FASTMAX_NVCC_EXTRA", "").split()


def sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def dev_sources() -> list[str]:
    d = os.path.join(CSRC, "dev")
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(CSRC, "dev"), os.path.join(HERE, "..", "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(f.encode() + b"\0" + fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_variant(name: str, extra: list[str], verbose: bool = False) -> str:
    """Developer helper: a second copy of the library under .variants/<name>/ compiled with extra nvcc
    flags (-DFM_TIMING, experiment switches); select it at run time with FASTMAX_LIB=<path>."""
    global LIB, OBJ, NVCC_FLAGS
    keep = (LIB, OBJ, NVCC_FLAGS)
    root = os.path.join(HERE, "..", ".variants", name)
    os.makedirs(root, exist_ok=True)
    LIB, OBJ, NVCC_FLAGS = os.path.join(root, "libfastmax_b200.so"), os.path.join(root, "build"), keep[2] + list(extra)
    try:
        return build(force=False, verbose=verbose)
    finally:
        LIB, OBJ, NVCC_FLAGS = keep


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ for sm_100a and link libfastmax_b200.so in-tree.
    Skips the work when the sources are unchanged since the last build."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    stamp = os.path.join(OBJ, "digest.txt")
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found and no up-to-date libfastmax_b200.so in tree")
    os.makedirs(OBJ, exist_ok=True)

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        return obj

    with cf.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        objs = list(ex.map(compile_one, sources()))
        dev_objs = list(ex.map(compile_one, dev_sources()))
    for out, oo in ((LIB, objs), (os.path.join(os.path.dirname(LIB), DEV_LIB_NAME), objs + dev_objs)):
        r = subprocess.run([nvcc, "-shared", "-o", out, *oo, "-lcudart"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as fh:
        fh.write(digest)
    return LIB


if __name__ == "__main__":
    import sys
    if "--variant" in sys.argv:      # python -m fastmax_experiments_b200.build --variant NAME -DFOO ...
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
