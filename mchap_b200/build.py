"""Build the C-ABI CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "libmchap_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",            # numba never contracts a*b+c; keep the reference's rounding
    "-Xcompiler", "-fPIC", "-shared",
    "-cudart", "static",      # the .so depends on the driver only, not on torch's runtime
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")
    return exe


def have_nvcc():
    return bool(shutil.which("nvcc")) or os.path.exists("/usr/local/cuda/bin/nvcc")


def sources():
    return sorted(
        os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))
    ) + [os.path.join(os.path.dirname(HERE), "include", "mchap_b200.h")]


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in sources())


PROFILE_LIB_PATH = os.path.join(LIB_DIR, "libmchap_b200_prof.so")


def build_profile(force=False):
    """The same library with -DMCHB_PROFILE (per-temperature cycle / event counters in the assemble
    kernel); loaded instead of the product build when MCHB_LIB points at it (profiles/phase_profile.py)."""
    if not force and os.path.exists(PROFILE_LIB_PATH) and \
            all(os.path.getmtime(s) <= os.path.getmtime(PROFILE_LIB_PATH) for s in sources()):
        return PROFILE_LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    subprocess.check_call([_nvcc()] + NVCC_FLAGS + ["-DMCHB_PROFILE", "-o", PROFILE_LIB_PATH,
                                                    os.path.join(CSRC, "libmchap_b200.cu")])
    return PROFILE_LIB_PATH


def build(force=False, verbose=False, extra_flags=()):
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    if os.environ.get("MCHB_ASM_MINBLOCKS"):
        extra_flags = list(extra_flags) + ["-DMCHB_ASM_MINBLOCKS=" + os.environ["MCHB_ASM_MINBLOCKS"]]
    cmd = [_nvcc()] + NVCC_FLAGS + list(extra_flags) + [
        "-o", LIB_PATH, os.path.join(CSRC, "libmchap_b200.cu"),
    ]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(LIB_PATH)
