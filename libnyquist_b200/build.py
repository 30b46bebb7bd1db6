"""In-tree build of the CUDA library (nvcc, sm_100a only).

    python -m libnyquist_b200.build

produces libnyquist_b200/lib/libnq_celt_b200.so next to the sources; the file
is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libnq_celt_b200.so")

# (source, extra flags): celt_synth_kernels.cu is compiled once per group of kernel instantiations
# (-DNQ_PART, see the top of that file) so the variants build in parallel.
UNITS = [("celt_synth_kernels.cu", ["-DNQ_PART=%d" % k], "celt_synth_kernels.part%d.o" % k) for k in range(4)] + [
    ("celt_post_kernels.cu", [], "celt_post_kernels.o"),
    ("celt_synth_api.cu", [], "celt_synth_api.o"),
    ("celt_frame_sink.cu", [], "celt_frame_sink.o"),
]
SOURCES = sorted({u[0] for u in UNITS})
HEADERS = ["celt_synth_kernels.cuh", "celt_fft_codelets.cuh", "celt_consts.cuh",
           os.path.join("..", "..", "include", "nq_celt_synth.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",   # B200 only: no PTX for other archs, no fallback
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
]


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(LIB_DIR, exist_ok=True)
    objdir = os.path.join(PKG, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")

    def compile_unit(unit):
        src, extra, obj = unit
        cmd = [nvcc, *NVCC_FLAGS, *extra, *os.environ.get("NQ_EXTRA_NVCC_FLAGS", "").split(), "-c", os.path.join(CSRC, src),
               "-o", os.path.join(objdir, obj)]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        return " ".join(cmd), r.returncode, r.stdout

    with ThreadPoolExecutor(max_workers=min(len(UNITS), os.cpu_count() or 2)) as ex:
        results = list(ex.map(compile_unit, UNITS))
    for cmd, rc, out in results:
        if verbose or rc != 0:
            print(cmd)
            print(out, end="")
        if rc != 0:
            raise subprocess.CalledProcessError(rc, cmd)
    link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-o", LIB_PATH,
            *[os.path.join(objdir, u[2]) for u in UNITS]]
    if verbose:
        print(" ".join(link))
    subprocess.run(link, check=True)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
