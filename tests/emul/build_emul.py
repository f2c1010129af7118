"""Build the TEST-ONLY sequential emulation of the CUDA kernel logic.

The product kernels (pve_mcc_for_unsignalized_intersection_b200/csrc/scene_step.cuh) are written
as barrier-separated phases.  Compiling the same source with g++ and -DPVE_HOST_EMULATION runs
every phase as a loop over thread ids on host memory.  This lets the CPU-only test tier
(`pytest -m "not gpu"`) check the kernel LOGIC -- the order-free reformulation of the sequential
reference -- against the oracle without a GPU.  It is never loaded by the product package
(BatchedScene refuses a non-CUDA backend unless a test passes the library explicitly), and it is
not a fallback: it is roughly as slow as one CPU thread gets.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "pve_mcc_for_unsignalized_intersection_b200", "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB = os.path.join(HERE, "_build", "libpve_emul.so")
SOURCES = [os.path.join(CSRC, "pve_mcc.cu"), os.path.join(CSRC, "scene_step.cuh"), os.path.join(CSRC, "scene_step4.cuh"),
           os.path.join(INCLUDE, "pve_mcc.h")]


def build_emul(force=False):
    if (not force and os.path.exists(LIB)
            and os.path.getmtime(LIB) >= max(os.path.getmtime(f) for f in SOURCES)):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = ["/usr/bin/g++", "-x", "c++", "-std=c++17", "-O2", "-g", "-fPIC", "-shared", "-ffp-contract=off", "-fno-strict-aliasing",
           "-DPVE_HOST_EMULATION", "-Wall", "-Wno-unknown-pragmas", "-I", INCLUDE, "-I", CSRC,
           SOURCES[0], "-o", LIB, "-lm"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + res.stdout)
    return LIB


if __name__ == "__main__":
    print(build_emul(force=True))
