"""Compile libep_b200.so in-tree with nvcc for sm_100a (no torch headers: the boundary is a plain C ABI)."""
import fcntl
import os
import subprocess
import sys

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB = os.path.join(CSRC, "libep_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    deps.append(os.path.join(os.path.dirname(os.path.dirname(CSRC)), "include", "ep_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    """Build the shared library if it is missing or older than its sources; returns its path."""
    if not force and not _stale():
        return LIB
    # one builder at a time (every rank of a torchrun launch calls this): build into a temporary file under a file
    # lock and rename it into place, so nobody ever dlopens a half-written library
    with open(os.path.join(CSRC, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale():                 # another process built it while we waited
                return LIB
            nvcc = os.environ.get("NVCC", "nvcc")
            tmp = f"{LIB}.tmp.{os.getpid()}"
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + _sources() + ["-o", tmp]
            r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed building libep_b200.so")
            os.replace(tmp, LIB)
            if verbose:
                sys.stderr.write(r.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
