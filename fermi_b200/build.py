"""Builds libfermi_b200.so (in-tree, sm_100a only) with nvcc.

    python -m fermi_b200.build [--force]

The library lands in fermi_b200/lib/ so that it travels with the source tree (gpurun snapshot) and is
found by fermi_b200._lib without any JIT cache.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libfermi_b200.so")

SOURCES = ["fmg_cuda.cu", "overlap.cu", "unitig_gpu.cu", "rld_enc.cu", "occ_build.cu", "build_bwt.cu", "bcr.cu", "ec.cu", "fmd_host.cpp", "occ_build_host.cpp", "unitig_host.cpp", "synth.cpp"]
HEADERS = ["cli_main.cpp", "fmd_device.cuh", "fmd_overlap.cuh", "fmd_host.hpp", "occ_layout.hpp", "fmg_internal.hpp", "ov_records.hpp", "dev_pool.hpp", "../../include/fermi_b200.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--expt-relaxed-constexpr", "--extended-lambda",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unknown-pragmas",
    "-Xptxas", "-v",
    "-shared",
]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    """Compile every CUDA/C++ source of the package into one shared library for sm_100a."""
    if not force and not _stale():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB] + [os.path.join(CSRC, f) for f in SOURCES] + ["-lpthread"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = os.path.join(LIBDIR, "build.log")
    with open(log, "w") as fh:
        fh.write(" ".join(cmd) + "\n" + res.stdout)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed, see " + log)
    # the command-line front end (fermi's command surface), host code linked against the library
    cli = [os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-Wall", "-o", os.path.join(HERE, "bin", "fermi-b200"),
           os.path.join(CSRC, "cli_main.cpp"), "-L" + LIBDIR, "-lfermi_b200", "-Wl,-rpath,$ORIGIN/../lib", "-lz"]
    os.makedirs(os.path.join(HERE, "bin"), exist_ok=True)
    res = subprocess.run(cli, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(log, "a") as fh:
        fh.write(" ".join(cli) + "\n" + res.stdout)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError("building the fermi-b200 front end failed, see " + log)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
