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

SOURCES = ["fmg_cuda.cu", "overlap.cu", "unitig_gpu.cu", "rld_enc.cu", "occ_build.cu", "build_bwt.cu", "bcr.cu", "ec.cu", "merge.cu", "contrast.cu", "rldx.cu", "fmd_host.cpp", "occ_build_host.cpp", "unitig_host.cpp", "synth.cpp", "bcr_compat.cpp"]
HEADERS = ["cli_main.cpp", "fmd_device.cuh", "fmd_overlap.cuh", "fmd_host.hpp", "occ_layout.hpp", "fmg_internal.hpp", "ov_records.hpp", "dev_pool.hpp", "bcr_tile.cuh", "../../include/fermi_b200.h"]

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


def _compile_one(nvcc, src, obj, log_lines):
    """one translation unit -> object (no relocatable device code: no device call crosses a file)"""
    cmd = [nvcc] + [f for f in NVCC_FLAGS if f != "-shared"] + ["-c", "-o", obj, src]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log_lines.append(" ".join(cmd) + "\n" + res.stdout)
    return res.returncode, res.stdout


def build_library(force=False, verbose=False):
    """Compile every CUDA/C++ source of the package for sm_100a (one object per source, stale ones only, in parallel)
    and link them into one shared library."""
    if not force and not _stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(os.path.dirname(HERE), "build", "obj")
    os.makedirs(objdir, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    hdr_t = max(os.path.getmtime(os.path.join(CSRC, f)) for f in HEADERS if not f.endswith(".cpp"))
    hdr_t = max(hdr_t, os.path.getmtime(os.path.abspath(__file__)))
    jobs, objs, logs = [], [], []
    for f in SOURCES:
        src, obj = os.path.join(CSRC, f), os.path.join(objdir, f + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t):
            jobs.append((src, obj))
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(lambda j: _compile_one(nvcc, j[0], j[1], logs), jobs))
    log = os.path.join(LIBDIR, "build.log")
    with open(log, "w" if force or len(jobs) == len(SOURCES) else "a") as fh:
        fh.write("".join(logs))
    for rc, out in results:
        if verbose or rc != 0:
            sys.stderr.write(out)
        if rc != 0:
            raise RuntimeError("nvcc failed, see " + log)
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs + ["-lpthread"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(log, "a") as fh:
        fh.write(" ".join(cmd) + "\n" + res.stdout)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError("linking libfermi_b200.so failed, see " + log)
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
