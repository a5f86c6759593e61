"""Build the UNMODIFIED reference extension ``PB_lib`` into oracle/_ref/ — TEST INFRASTRUCTURE ONLY.

Sources are compiled where they lie under /root/reference/lib/PB_lib/src (nothing is copied into the
repo); outputs go to oracle/_ref/ only (git-ignored, but shipped to the GPU box by gpurun).  The
reference's own setup.py (lib/PB_lib/setup.py:6-10) is not run; this recipe issues the same three
translation units directly:

    src/PB_lib_api.cpp  src/PB_lib.cpp  src/cuda.cu      (unity build, lib/PB_lib/src/cuda.cu:2-8)

The only deviation is a forced ``-include thrust/sort.h`` (+ tuple / zip_iterator) for nvcc: CUDA 12.9's
thrust no longer pulls ``sort_by_key`` in transitively (binary.cu:68 would not compile otherwise).
No source edits.  The reference has no CPU path: the resulting module needs a GPU to *run*, so it is
exercised only by the ``-m gpu`` cross-check tests and by ``bench.py --impl reference``.
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_SRC = "/root/reference/lib/PB_lib/src"


def so_path() -> str:
    return os.path.join(OUT, "PB_lib" + sysconfig.get_config_var("EXT_SUFFIX"))


REF_WRAPPER = "/root/reference/lib/PB_lib/torch_io/pbnet_ops.py"


def wrapper_pyc() -> str:
    return os.path.join(OUT, "ref_pbnet_ops.pyc.bin")


def build_wrapper(force: bool = False) -> str | None:
    """Byte-compiles the UNMODIFIED reference wrapper (lib/PB_lib/torch_io/pbnet_ops.py) into oracle/_ref/ so that the
    zero-edit drop-in test can execute it on the GPU box, where /root/reference does not exist.  A compiled output like
    the .so: no reference source enters the repo."""
    pyc = wrapper_pyc()
    if not os.path.exists(REF_WRAPPER):
        return pyc if os.path.exists(pyc) else None
    if force or not os.path.exists(pyc):
        import py_compile
        os.makedirs(OUT, exist_ok=True)
        py_compile.compile(REF_WRAPPER, cfile=pyc, dfile="lib/PB_lib/torch_io/pbnet_ops.py", doraise=True)
    return pyc


def build(force: bool = False) -> str | None:
    """Returns the path of the built module, or None when /root/reference is absent (GPU box)."""
    build_wrapper(force)
    so = so_path()
    if not os.path.isdir(REF_SRC):
        return so if os.path.exists(so) else None
    if os.path.exists(so) and not force:
        return so
    from torch.utils import cpp_extension as ce

    os.makedirs(OUT, exist_ok=True)
    inc = []
    for p in ce.include_paths("cuda") + [sysconfig.get_paths()["include"], REF_SRC]:
        inc += ["-I", p]
    common = ["-DTORCH_EXTENSION_NAME=PB_lib", "-DTORCH_API_INCLUDE_EXTENSION_H", "-D_GLIBCXX_USE_CXX11_ABI=1",
              "-std=c++17"]
    objs = []
    for src in ("PB_lib_api.cpp", "PB_lib.cpp"):
        o = os.path.join(OUT, src + ".o")
        subprocess.check_call(["g++", "-O2", "-g", "-fPIC", "-w", *common, *inc, "-c",
                               os.path.join(REF_SRC, src), "-o", o])
        objs.append(o)
    o = os.path.join(OUT, "cuda.cu.o")
    subprocess.check_call(["nvcc", "-O2", "-w", "-gencode", "arch=compute_100a,code=sm_100a",
                           "-include", "thrust/sort.h", "-include", "thrust/tuple.h",
                           "-include", "thrust/iterator/zip_iterator.h",
                           "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", *common, *inc, "-c",
                           os.path.join(REF_SRC, "cuda.cu"), "-o", o])
    objs.append(o)
    libdirs = []
    for p in ce.library_paths("cuda"):
        libdirs += ["-L", p, "-Wl,-rpath," + p]
    subprocess.check_call(["g++", "-shared", *objs, *libdirs, "-lc10", "-ltorch_cpu", "-ltorch",
                           "-ltorch_python", "-lc10_cuda", "-ltorch_cuda", "-lcudart", "-o", so])
    for o in objs:
        os.remove(o)
    return so


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
