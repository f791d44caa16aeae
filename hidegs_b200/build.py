"""In-tree build of the CUDA library (sm_100a) with plain nvcc.

`python -m hidegs_b200.build` compiles hidegs_b200/csrc/*.cu into
hidegs_b200/libhidegs_b200.so.  nvcc cross-compiles without a GPU, so this runs
in the CPU-only build container; the .so travels to the GPU box with the tree.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libhidegs_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
          "-Xptxas", "-v", "--expt-relaxed-constexpr"]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp")))  # .cpp: host compiler only


def _digest(paths):
    h = hashlib.sha256()
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(ARCH + CFLAGS).encode())
    return h.hexdigest()


def _compile(src, verbose):
    obj = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
    stamp = obj + ".sha"
    deps = [os.path.join(CSRC, src)] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    inc = os.path.join(HERE, "..", "include")
    deps += [os.path.join(inc, f) for f in sorted(os.listdir(inc)) if f.endswith(".h")]
    dig = _digest(deps)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj, ""
    cmd = [NVCC] + ARCH + CFLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, res.stdout, res.stderr))
    with open(stamp, "w") as f:
        f.write(dig)
    return obj, res.stderr if verbose else ""


EXT_SRC = os.path.join(HERE, "csrc_ext", "raster_ext.cpp")
EXT_NAME = "_hgC"


def ext_path():
    import sysconfig
    return os.path.join(HERE, "diff_gaussian_rasterization", EXT_NAME + sysconfig.get_config_var("EXT_SUFFIX"))


def build_extension(verbose=False, force=False):
    """The thin torch C++ extension of the rasterizer operators (csrc_ext/raster_ext.cpp -> diff_gaussian_rasterization/
    _hgC*.so): host code only (g++), linked against libhidegs_b200.so next to it (rpath $ORIGIN/..) and torch's libraries
    (already loaded when the module is imported)."""
    import sysconfig
    import torch
    from torch.utils import cpp_extension as ce
    out = ext_path()
    inc = os.path.join(HERE, "..", "include")
    deps = [EXT_SRC, LIB] + [os.path.join(inc, f) for f in sorted(os.listdir(inc)) if f.endswith(".h")]
    stamp = out + ".sha"
    dig = _digest(deps[:1] + deps[2:]) + torch.__version__
    if not force and os.path.exists(out) and os.path.exists(stamp) and open(stamp).read() == dig:
        return out
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = [os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden",
           "-DTORCH_EXTENSION_NAME=" + EXT_NAME, "-DTORCH_API_INCLUDE_EXTENSION_H",
           "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    cmd += ["-I" + p for p in ce.include_paths()] + ["-I" + sysconfig.get_paths()["include"], "-I" + cuda_home + "/include"]
    cmd += [EXT_SRC, "-o", out]
    cmd += ["-L" + p for p in ce.library_paths()] + ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch",
                                                     "-ltorch_python"]
    cmd += ["-L" + HERE, "-l:libhidegs_b200.so", "-Wl,-rpath,$ORIGIN/.."]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building the torch extension failed:\n%s\n%s" % (res.stdout, res.stderr))
    with open(stamp, "w") as f:
        f.write(dig)
    if verbose:
        sys.stderr.write(res.stderr)
    return out


def build(verbose=False, force=False):
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    srcs = sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile(s, verbose), srcs))
    objs = [r[0] for r in results]
    log = "".join(r[1] for r in results)
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart_static", "-Xcompiler", "-fPIC"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (res.stdout, res.stderr))
    if verbose and log:
        sys.stderr.write(log)
    build_extension(verbose=verbose, force=force)
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
