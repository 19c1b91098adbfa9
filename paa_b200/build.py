"""In-tree build of libpaa_b200.so (nvcc, sm_100a only).

The library has no torch / pybind dependency: plain CUDA runtime behind a C ABI
(include/paa_b200.h).  ``python -m paa_b200.build`` or ``__graft_entry__.build()`` compile it next
to the sources; the built .so travels to the GPU box with the repo snapshot.
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libpaa_b200.so")
STAMP_PATH = LIB_PATH + ".stamp"
SOURCES = ["abi.cu", "assign.cu", "loss.cu", "post.cu", "aux.cu", "atss.cu", "retina.cu", "fcos.cu", "rpn.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xptxas=-v", "-Xcompiler", "-fPIC", "-shared",
              "-cudart", "shared"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libpaa_b200.so cannot be built")


def _fingerprint():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for name in sorted(os.listdir(root)):
            if name.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, name), "rb") as f:
                    h.update(name.encode())
                    h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compiles the library if the sources changed since the last build.  Returns the .so path."""
    fp = _fingerprint()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(STAMP_PATH):
        with open(STAMP_PATH) as f:
            if f.read().strip() == fp:
                return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB_PATH]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed (exit %d):\n%s" % (proc.returncode, proc.stdout[-4000:]))
    with open(LIB_PATH + ".ptxas.log", "w") as f:
        f.write(proc.stdout)
    with open(STAMP_PATH, "w") as f:
        f.write(fp)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
