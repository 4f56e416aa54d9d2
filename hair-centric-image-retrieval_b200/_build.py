"""In-tree build of libhcir_b200.so (nvcc, sm_100a only).  No torch C++ API involved: the
library is a plain C-ABI shared object (include/hcir_b200.h) loaded with ctypes."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libhcir_b200.so")
STAMP = os.path.join(HERE, "build", "stamp.txt")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "--expt-relaxed-constexpr",
] + os.environ.get("HCIR_NVCC_EXTRA", "").split()   # A/B builds on the GPU box (e.g. -DHCIR_DOT_GROUP=2)


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (needed to build libhcir_b200.so for sm_100a)")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest() -> str:
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + ["../../include/hcir_b200.h"]
    for f in files:
        p = os.path.normpath(os.path.join(CSRC, f))
        if os.path.isfile(p):
            h.update(f.encode())
            h.update(open(p, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    try:
        return os.path.exists(LIB_PATH) and open(STAMP).read().strip() == _digest()
    except OSError:
        return False


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile every csrc/*.cu for sm_100a and link libhcir_b200.so in-tree."""
    if not force and is_current():
        return LIB_PATH
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    # One builder at a time: under torchrun every rank of a fresh checkout gets here at once.  The
    # others wait on the lock, find the stamp current and return; the .so is linked to a temporary
    # name and renamed into place, so nobody ever maps a half-written library.
    import fcntl
    with open(os.path.join(objdir, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and is_current():
                return LIB_PATH
            return _build_locked(objdir, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(objdir: str, verbose: bool) -> str:
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=8) as ex:
        results = list(ex.map(compile_one, sources()))
    if verbose:
        for _, log in results:
            print(log)
    objs = [o for o, _ in results]
    tmp = LIB_PATH + f".tmp{os.getpid()}"
    r = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp, *objs],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB_PATH)
    with open(STAMP + ".tmp", "w") as f:
        f.write(_digest())
    os.replace(STAMP + ".tmp", STAMP)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
