"""In-tree build of libou_b200.so with nvcc for sm_100a (cross-compiles without a GPU).

    python -m open_universe_b200.build [--force] [--verbose]

The shared library lands next to its sources (``csrc/libou_b200.so``) so that it travels with the
repo snapshot to the GPU box and shows up as an in-tree native library when loaded.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

CSRC = Path(__file__).resolve().parent / "csrc"
# OU_ACT_BF16=1 builds (and engine.lib loads) the bf16 storage policy as a separate library for A/B
# runs; the product default is fp16 storage (csrc/common.cuh)
BF16_VARIANT = os.environ.get("OU_ACT_BF16", "0") not in ("", "0")
LIB = CSRC / ("libou_b200_bf16.so" if BF16_VARIANT else "libou_b200.so")
SOURCES = ["api.cu", "conv.cu", "conv_tc.cu", "conv_trunk.cu", "gru.cu", "signal.cu", "mel.cu", "snake.cu", "metrics.cu", "plan.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not LIB.exists():
        return True
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + \
        [CSRC.parent.parent / "include" / "ou_b200.h"]
    return any(d.stat().st_mtime > LIB.stat().st_mtime for d in deps)


def build(force=False, verbose=False):
    """Compile and link in-tree.  Serialised across processes with a file lock: the ranks of a torchrun
    launch all call ``lib.load()`` at start-up and must not run nvcc into the same files at once."""
    if not force and not needs_build():
        return LIB
    import fcntl
    with open(CSRC / ".build.lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():      # another rank built it while we waited
                return LIB
            return _build_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose):
    nvcc = _nvcc()
    objs = []

    def compile_one(src):
        obj = CSRC / (Path(src).stem + (".bf16.o" if BF16_VARIANT else ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *(["-DOU_ACT_BF16"] if BF16_VARIANT else []), "-c", str(CSRC / src),
               "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    tmp = LIB.with_suffix(".so.tmp")                 # link aside, then rename: readers never see half a file
    cmd = [nvcc, "-shared", "-o", str(tmp), *map(str, objs), "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
