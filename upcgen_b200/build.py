"""Builds libupcgpu.so (the C-ABI library, include/upcgpu.h) in-tree with nvcc for sm_100a.

    python -m upcgen_b200.build [--force]

nvcc cross-compiles without a GPU; the resulting .so travels to the GPU box with the repo
snapshot (it is git-ignored, not gpurun-ignored).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libupcgpu.so")
SOURCES = ["upc_capi.cu", "upc_tables.cu", "upc_lumi.cu", "upc_fold.cu", "upc_events.cu", "upc_group.cu", "upc_elem_capi.cpp",
           "../host/UpcTwoPhotonDilep.cpp", "../host/UpcTwoPhotonALP.cpp", "../host/UpcTwoPhotonTabulated.cpp",
           "../host/UpcRootHist.cpp", "../host/UpcRootFile.cpp", "../host/UpcLz4.cpp"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--fmad=true", "-Xptxas", "-v", "-ccbin", "/usr/bin/g++"]


def _newest_src():
    t = 0.0
    for root in (CSRC, os.path.join(HERE, "host"), os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(root):
            if f.endswith((".cu", ".cuh", ".cpp", ".h")):   # sources only: host/ also holds libupcgen_host.so and upcgen
                t = max(t, os.path.getmtime(os.path.join(root, f)))
    return t


def build(force=False, verbose=False):
    if not force and os.path.exists(SO) and os.path.getmtime(SO) >= _newest_src():
        return SO
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for s in SOURCES:
        o = os.path.join(HERE, "build", os.path.basename(s).rsplit(".", 1)[0] + ".o")
        objs.append(o)
        cmd = [NVCC, *FLAGS, "-Xcompiler", "-ffp-contract=off", "-c", os.path.join(CSRC, s), "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {s} ====\n{out}")
        failed |= p.returncode != 0
    with open(os.path.join(HERE, "build", "nvcc.log"), "w") as fh:
        fh.write("\n".join(log))
    if failed or verbose:
        sys.stderr.write("\n".join(log))
    if failed:
        raise RuntimeError("nvcc failed, see upcgen_b200/build/nvcc.log")
    subprocess.check_call([NVCC, "-shared", "-Wno-deprecated-gpu-targets", "-o", SO, *objs, "-lcudart", "-lz", "-ldl", "-lpthread", "-ccbin", "/usr/bin/g++"])
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
