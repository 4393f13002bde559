"""Builds libnlo_cuda.so (sm_100a) in-tree with nvcc.  No GPU is needed to compile."""
import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_PKG, "csrc")
LIB_PATH = os.path.join(_PKG, "libnlo_cuda.so")
SOURCES = ["nlo_kernels.cu", "nlo_api.cu", "nlo_ingest.cu", "nlo_map.cu", "nlo_multi.cu"]
HEADERS = ["nlo_internal.h", "nlo_host.h", "nlo_device.cuh", os.path.join("..", "..", "include", "nlo_cuda.h")]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(_CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build_cuda(force=False, verbose=False):
    """Compile csrc/*.cu into nonlinear_optimizer_for_slam_b200/libnlo_cuda.so."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(_CSRC, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(_CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    for cmd, proc in procs:
        out, _ = proc.communicate()
        if verbose or proc.returncode != 0:
            print(out.decode())
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    link = [nvcc, "-shared", "-Wno-deprecated-gpu-targets", "-o", LIB_PATH] + objs + ["-lcudart", "-ldl", "-lpthread"]
    subprocess.run(link, check=True)
    return LIB_PATH


CXX_TEST_SRC = os.path.join(_PKG, "cxx", "tests", "simple_optimization_test.cc")
CXX_TEST_BIN = os.path.join(_PKG, "cxx", "tests", "simple_optimization_test")


def build_cxx_example(force=False):
    """Compile the C++ drop-in example / test against libnlo_cuda.so (g++, no CUDA headers)."""
    build_cuda()
    if (not force and os.path.exists(CXX_TEST_BIN)
            and os.path.getmtime(CXX_TEST_BIN) >= os.path.getmtime(CXX_TEST_SRC)
            and os.path.getmtime(CXX_TEST_BIN) >= os.path.getmtime(LIB_PATH)):
        return CXX_TEST_BIN
    cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-Wextra",
           "-I", os.path.join(_PKG, "..", "include"), "-I", os.path.join(_PKG, "cxx"),
           CXX_TEST_SRC, "-o", CXX_TEST_BIN, "-L", _PKG, "-lnlo_cuda",
           "-Wl,-rpath," + _PKG, "-Wl,-rpath,$ORIGIN/../.."]
    subprocess.run(cmd, check=True)
    return CXX_TEST_BIN


CXX_BENCH_SRC = os.path.join(_PKG, "cxx", "bench", "dropin_bench.cc")
CXX_BENCH_BIN = os.path.join(_PKG, "cxx", "bench", "dropin_bench")


def build_cxx_bench(force=False):
    """Compile the drop-in end-to-end benchmark (times ...Cuda::Solve on std::vector<Correspondence>)."""
    build_cuda()
    if (not force and os.path.exists(CXX_BENCH_BIN)
            and os.path.getmtime(CXX_BENCH_BIN) >= os.path.getmtime(CXX_BENCH_SRC)
            and os.path.getmtime(CXX_BENCH_BIN) >= os.path.getmtime(LIB_PATH)):
        return CXX_BENCH_BIN
    cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-Wextra", "-pthread",
           "-I", os.path.join(_PKG, "..", "include"), "-I", os.path.join(_PKG, "cxx"),
           CXX_BENCH_SRC, "-o", CXX_BENCH_BIN, "-L", _PKG, "-lnlo_cuda",
           "-Wl,-rpath," + _PKG, "-Wl,-rpath,$ORIGIN/../.."]
    subprocess.run(cmd, check=True)
    return CXX_BENCH_BIN


if __name__ == "__main__":
    import sys
    print(build_cuda(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
