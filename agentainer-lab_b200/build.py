"""In-tree build of the CUDA library for sm_100a (nvcc cross-compiles without a GPU)."""
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIBDIR = os.path.join(_HERE, "lib")
LIBNAME = "libagentainer_b200.so"
SOURCES = ["agr_kernels.cu", "agr_k1_tma.cu", "agr_k1_var.cu", "agr_k4.cu", "agr_k5_json.cu", "agr_json_host.cpp", "agr_engine.cu"]
HEADERS = ["agr_common.h", "agr_synth.h", "agr_kernels.cuh", "agr_device.cuh", os.path.join("..", "..", "include", "agentainer_gpu.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def lib_path() -> str:
    return os.path.join(LIBDIR, LIBNAME)


def _stale() -> bool:
    out = lib_path()
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build_native(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu into lib/libagentainer_b200.so.  Returns the library path."""
    if not force and not _stale():
        return lib_path()
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libagentainer_b200.so (and there is no CPU fallback)")
    os.makedirs(LIBDIR, exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", lib_path(), "-lcudart", "-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return lib_path()


HOST = os.path.join(_HERE, "host")


def build_host(force: bool = False) -> str:
    """Compile the C++ mirror of requests.Manager / ReplayWorker and its threaded driver (host/test_host)."""
    out = os.path.join(HOST, "test_host")
    srcs = [os.path.join(HOST, f) for f in ("requests.cpp", "test_host.cpp")]
    deps = srcs + [os.path.join(HOST, "requests.hpp"), lib_path()]
    if not force and os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps if os.path.exists(d)):
        return out
    gxx = shutil.which("g++")
    if not gxx:
        raise RuntimeError("g++ not found")
    cmd = [gxx, "-O2", "-std=c++17", "-Wall", "-I" + os.path.join(_HERE, "..", "include")] + srcs + [
        "-o", out, "-L" + LIBDIR, "-lagentainer_b200", "-Wl,-rpath,$ORIGIN/../lib", "-lpthread"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + res.stdout + res.stderr)
    return out
