"""Builds libtracefly_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python tracefly-ga-uav-pathplanning_b200/build.py [--force] [--verbose]

The .so lands in tracefly-ga-uav-pathplanning_b200/_build/ (git-ignored, travels with gpurun).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# TF_BUILD_DIR / TF_NVCC_EXTRA: A/B builds of kernel variants (e.g. -DTF_DUO_XF1_CTAS=2) next to the product
# library; tracefly_b200._lib loads such a variant when TRACEFLY_B200_LIB points at it.  TF_NO_LINEINFO=1 drops -lineinfo
# (smaller A/B libraries).
OUT_DIR = os.environ.get("TF_BUILD_DIR") or os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libtracefly_b200.so")
SOURCES = ["tf_env.cu", "tf_fitness.cu", "tf_rank.cu", "tf_ga.cu", "tf_io.cu", "tf_tape.cu"]
HEADERS = ["tf_common.cuh", "tf_fitness_dev.cuh", "tf_ga_dev.cuh", "tf_stream_dev.cuh", "tf_stream_host.cuh",
           "tf_fitness_kernels.cuh", "tf_generation_kernels.cuh", "tf_inst.cu", "../../include/tracefly_b200.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--fmad=false",            # every FMA is explicit: results must match the fixed-order oracle bit for bit
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


# tf_inst.cu is compiled once per (kernel family, waypoints-per-lane): the template instantiations
# (hundreds of kernels) then build in parallel instead of inside two giant translation units.
INSTANCES = [(kind, nw) for kind in (0, 1) for nw in (1, 2, 4, 8)] + [(kind, nw) for kind in (2, 3) for nw in (1, 2, 4, 8)]
MAX_JOBS = max(1, min(8, os.cpu_count() or 1))


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    jobs = []
    for src in SOURCES:
        obj = os.path.join(OUT_DIR, src.replace(".cu", ".o"))
        jobs.append((src, obj, []))
    for kind, nw in INSTANCES:
        obj = os.path.join(OUT_DIR, f"tf_inst_k{kind}_nw{nw}.o")
        jobs.append(("tf_inst.cu", obj, [f"-DTF_INST_KIND={kind}", f"-DTF_INST_NW={nw}"]))
    # longest jobs first (the stream / generation instantiations), at most MAX_JOBS compilers at a time
    jobs.sort(key=lambda j: (j[0] != "tf_inst.cu", j[1]))
    objs = [j[1] for j in jobs]
    running, pending = [], list(jobs)
    while pending or running:
        while pending and len(running) < MAX_JOBS:
            src, obj, defs = pending.pop(0)
            cmd = [nvcc()] + [f for f in NVCC_FLAGS if f != "-lineinfo" or not os.environ.get("TF_NO_LINEINFO")] + os.environ.get("TF_NVCC_EXTRA", "").split() + defs + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
            running.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        src, obj, proc = running.pop(0)
        out, _ = proc.communicate()
        if verbose or proc.returncode != 0:
            sys.stderr.write(out)
        if proc.returncode != 0:
            for _, _, p in running:
                p.kill()
            raise RuntimeError(f"nvcc failed on {src} -> {os.path.basename(obj)}")
    cmd = [nvcc(), "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
