"""
In-tree build of the CUDA library (nvcc cross-compiles sm_100a without a GPU):

    python -m rajepy_b200.build            # -> rajepy_b200/lib/librajepy_b200.so

The .so is git-ignored but travels with the repo snapshot to the GPU box.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.environ.get("RAJEPY_B200_LIB") or os.path.join(LIBDIR, "librajepy_b200.so")
SOURCES = ["rjp_api.cu", "rjp_fill.cu", "rjp_integrate.cu", "rjp_host.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
              "-std=c++17", "--shared", "-Xcompiler", "-fPIC,-pthread", "-Xptxas", "-v"]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found")
    return exe


def have_nvcc():
    try:
        _nvcc()
        return True
    except RuntimeError:
        return False


def source_hash():
    """Content hash of everything the library is built from (mtimes do not survive the copy
    to the GPU box)."""
    import hashlib
    h = hashlib.sha256()
    deps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC))
    deps.append(os.path.join(os.path.dirname(HERE), "include", "rajepy_b200.h"))
    for d in deps:
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS + os.environ.get("RAJEPY_B200_NVCC_EXTRA", "").split()).encode())
    return h.hexdigest()


def needs_build():
    if not os.path.exists(LIB):
        return True
    try:
        with open(LIB + ".hash") as f:
            return f.read().strip() != source_hash()
    except OSError:
        return True


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    extra = os.environ.get("RAJEPY_B200_NVCC_EXTRA", "").split()  # kernel-variant experiments
    cmd = [_nvcc()] + NVCC_FLAGS + extra + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(LIBDIR, "build.log"), "wt") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed")
    with open(LIB + ".hash", "wt") as f:
        f.write(source_hash())
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
