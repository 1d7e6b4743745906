"""Build librendertoy_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m rendertoy_b200.build [--force] [--verbose]

Flags that matter:
  -gencode arch=compute_100a,code=sm_100a   B200 only, no PTX fallback for other parts
  -fmad=false                               no FMA contraction: float arithmetic is evaluated as written,
                                            which is what makes the raster path bit-exact against the oracle
  -lineinfo                                 so `ncu --import-source on` maps SASS to these files
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(HERE, "build")
LIB_PATH = os.path.join(HERE, "librendertoy_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-O2,-fno-fast-math",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: librendertoy_b200.so cannot be built (there is no CPU fallback)")


def _host_compiler_args():
    # the image's $CC/$CXX wrapper lacks some specs; nvcc is happiest with the distro g++
    return ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "rendertoy_b200.h"))
    hdrs.append(os.path.abspath(__file__))
    return max(os.path.getmtime(h) for h in hdrs if os.path.exists(h))


def build_native(force=False, verbose=False):
    """Compile every csrc/*.cu and link the shared library; returns its path."""
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc, hdr_m = _nvcc(), _deps_mtime()
    jobs = []
    for src in sources():
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_m):
            cmd = [nvcc, *NVCC_FLAGS, *_host_compiler_args(), "-c", src, "-o", obj]
            if verbose:
                cmd[1:1] = ["-Xptxas", "-v"]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0 or verbose:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for " + cmd[-3])

    with ThreadPoolExecutor(max_workers=4) as ex:
        list(ex.map(run, jobs))
    objs = [os.path.join(OBJ_DIR, os.path.basename(s)[:-3] + ".o") for s in sources()]
    if jobs or force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(o) > os.path.getmtime(LIB_PATH) for o in objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", *_host_compiler_args(), "-o", LIB_PATH, *objs,
               "-cudart", "static", "-ldl"]
        run(cmd)
    return LIB_PATH


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
