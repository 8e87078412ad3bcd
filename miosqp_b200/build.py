"""
Builds libbqp.so (the C-ABI engine of include/bqp.h) IN-TREE with nvcc for sm_100a.

    python -m miosqp_b200.build [--force]

The shared object lands next to this file (miosqp_b200/libbqp.so); it is git-ignored but
travels to the GPU box with the gpurun snapshot.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# A/B builds of kernel variants: BQP_BUILD_DEFS="-DBQP_P1_SETS=2 ..." BQP_LIB_SUFFIX=_v2 -> libbqp_v2.so (same digest guard)
EXTRA_DEFS = os.environ.get("BQP_BUILD_DEFS", "").split()
OUT = os.path.join(HERE, "libbqp%s.so" % os.environ.get("BQP_LIB_SUFFIX", ""))
SOURCES = ["bqp_setup.cpp", "bqp_bnb.cpp", "bqp_kernels.cu", "bqp_stream.cu", "bqp_panel.cu", "bqp_rows.cu", "bqp_grid.cu", "bqp_small.cu", "bqp_api.cu"]
HEADERS = ["bqp_internal.h", os.path.join("..", "..", "include", "bqp.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-O3,-Wall", "-shared", "-Xptxas", "-v"]


def _digest():
    h = hashlib.sha256()
    for d in [os.path.join(CSRC, s) for s in SOURCES + HEADERS]:
        with open(d, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS + EXTRA_DEFS).encode())
    return h.hexdigest()


def _stale():
    # content hash, not mtimes: the gpurun snapshot does not preserve timestamps
    if not os.path.exists(OUT) or not os.path.exists(OUT + ".sha256"):
        return True
    with open(OUT + ".sha256") as f:
        return f.read().strip() != _digest()


def build(force=False, verbose=False):
    """Compile when the sources changed.  Raises on failure (no fallback).  Safe when several processes call it at once
    (every torchrun rank does): one builds under a file lock into a temporary file and renames it into place, the
    others wait on the lock and then find the library fresh."""
    if not force and not _stale():
        return OUT
    import fcntl
    with open(OUT + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale():          # another process built it while we waited
                return OUT
            nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
            tmp = "%s.tmp.%d" % (OUT, os.getpid())
            cmd = [nvcc] + NVCC_FLAGS + EXTRA_DEFS + ["-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES]
            res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            log = os.path.join(HERE, "build%s.log" % os.environ.get("BQP_LIB_SUFFIX", ""))
            with open(log, "w") as f:
                f.write(" ".join(cmd) + "\n" + res.stdout)
            if verbose or res.returncode != 0:
                sys.stderr.write(res.stdout)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed building %s (see %s)" % (os.path.basename(OUT), log))
            os.replace(tmp, OUT)
            with open(OUT + ".sha256.tmp", "w") as f:
                f.write(_digest())
            os.replace(OUT + ".sha256.tmp", OUT + ".sha256")
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(OUT)
