"""Build recipe of libbarbu_hair.so (sm_100a only, in-tree so the .so travels with the snapshot)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libbarbu_hair.so")
SOURCES = ["hair_step.cu", "hair_stream.cu", "hair_wave.cu", "hair_gen.cu", "hair_tess.cu", "hair_state.cu", "hair_marschner.cu", "hair_capi.cu", "hair_group.cu", "hair_host.cc"]
HEADERS = ["hair_step.cuh", "hair_gen.cuh", "hair_math.cuh", "hair_collide.cuh", "hair_sim.cuh", os.path.join("..", "..", "include", "barbu_hair.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    # host side: one rounding per written fp operation (strand generation restates host arithmetic)
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-Wall",
    "-ccbin", "/usr/bin/g++",
    "-shared",
]


def nvcc_path():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libbarbu_hair.so can only be built with the CUDA toolkit")


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every CUDA source of the package into barbu_b200/lib/libbarbu_hair.so."""
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", LIB_PATH] + [os.path.join(CSRC, f) for f in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    return LIB_PATH


ADAPTOR_SRC = os.path.join(os.path.dirname(HERE), "tests", "cpp", "hair_adaptor_main.cc")
ADAPTOR_BIN = os.path.join(LIB_DIR, "hair_adaptor_main")


def build_cpp_adaptor_driver(force=False):
    """g++ -std=c++17 build of the C++ `Hair` adaptor driver (tests/cpp) against libbarbu_hair.so — host compiler only:
    the adaptor header needs neither CUDA nor GL headers."""
    hdrs = [os.path.join(os.path.dirname(HERE), "include", h) for h in ("barbu_hair.h", "barbu_hair.hpp")]
    if not force and os.path.exists(ADAPTOR_BIN) and all(os.path.getmtime(d) <= os.path.getmtime(ADAPTOR_BIN)
                                                          for d in hdrs + [ADAPTOR_SRC, LIB_PATH]):
        return ADAPTOR_BIN
    cmd = ["/usr/bin/g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-o", ADAPTOR_BIN, ADAPTOR_SRC,
           "-L" + LIB_DIR, "-lbarbu_hair", "-Wl,-rpath,$ORIGIN"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    return ADAPTOR_BIN


if __name__ == "__main__":
    print(build(force=True, verbose=True))
