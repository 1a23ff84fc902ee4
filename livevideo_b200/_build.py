"""In-tree build of the CUDA library (sm_100a).  `python -m livevideo_b200._build` or __graft_entry__.build()."""
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "liblivevideo.so")
SOURCES = ["lv_kernels.cu", "lv_zerocopy.cu", "lv_tma.cu", "lv_aux.cu", "lv_formats.cu", "lv_shard.cu", "lv_exchange.cu", "lv_capi.cu", "lv_synth.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: the CUDA library cannot be built")
    return p


def _deps():
    out = [os.path.join(PKG_DIR, "..", "include", "livevideo.h")]
    for f in os.listdir(CSRC):
        if f.endswith((".cu", ".cuh", ".h", ".cpp")):
            out.append(os.path.join(CSRC, f))
    return out


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into livevideo_b200/liblivevideo.so."""
    if not force and not needs_build():
        return LIB_PATH
    cmd = [nvcc_path(), *ARCH, "-O3", "-lineinfo", "-std=c++17", "-DLV_BUILDING_LIBRARY", "-shared", "-Xcompiler", "-fPIC,-fvisibility=hidden",
           "-Xptxas", "-v" if verbose else "-O3", "-cudart", "shared", "-ldl", "-o", LIB_PATH] + \
          [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building liblivevideo.so")
    if verbose:
        sys.stderr.write(r.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
