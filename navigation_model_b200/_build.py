"""In-tree build of the C-ABI library (nvcc, sm_100a).  Used by ``__graft_entry__.build()``.

The built ``libnavmodel_b200.so`` stays next to this file (git-ignored, but shipped to the GPU box by
``gpurun``).  nvcc cross-compiles without a GPU, so this runs in the CPU build container too.
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_NAME = "libnavmodel_b200.so"
LIB_PATH = os.path.join(HERE, LIB_NAME)

SOURCES = ["nm_error.cpp", "nm_mouse.cpp", "nm_stream.cpp", "nm_maze.cu", "nm_fit.cu", "nm_sim.cu", "nm_mb.cu", "nm_q.cu", "nm_dist.cu", "nm_session.cu", "nm_trace.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",                      # parity: no implicit FMA contraction on the device
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2,-Wall",
    "-Xptxas", "-v",
    "-shared",
]


def find_nvcc():
    cand = [os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"]
    for c in cand:
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def needs_rebuild():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "navmodel_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_native(force=False, verbose=False):
    """Compile every CUDA / C++ source of the package into ``libnavmodel_b200.so``."""
    if not force and not needs_rebuild():
        return LIB_PATH
    extra = os.environ.get("NM_NVCC_EXTRA", "").split()  # tuning experiments, e.g. -DNM_SIM_MIN_CTAS=4
    cmd = [find_nvcc()] + NVCC_FLAGS + extra + ["-x", "cu"] + sources() + ["-o", LIB_PATH + ".tmp"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(HERE, "build.log"), "w") as fh:
        fh.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + log[-4000:])
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    if verbose:
        print(log)
    return LIB_PATH


if __name__ == "__main__":
    print(build_native(force=True, verbose=True))
