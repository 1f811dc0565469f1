"""In-tree build of libtocvp.so (explicit nvcc, sm_100a only).  `python -m textocvp_b200.build`."""
from __future__ import annotations

import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtocvp.so")
OBJ = os.path.join(HERE, "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "--expt-relaxed-constexpr"]
# test-only native code (hardware probes): its own small library, NOT part of libtocvp.so or of include/tocvp.h
TEST_NATIVE = os.path.join(HERE, "..", "tests", "native")
TEST_LIB = os.path.join(TEST_NATIVE, "libtocvp_probe.so")


def _deps_mtime():
    hdrs = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    return max(os.path.getmtime(h) for h in hdrs)


def build(verbose: bool = False, force: bool = False, ptxas_v: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdr_t = _deps_mtime()
    jobs = []
    objs = []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_t):
            cmd = [NVCC, *FLAGS, "-c", s, "-o", o]
            if ptxas_v:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        for cmd, r in ex.map(run, jobs):
            if verbose or r.returncode != 0 or ptxas_v:
                sys.stderr.write(" ".join(cmd[-3:]) + "\n" + r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {cmd[-3]}")
    if jobs or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    build_test_native(hdr_t, force)
    return LIB


def build_test_native(hdr_t: float, force: bool = False) -> None:
    """tests/native/*.cu -> tests/native/libtocvp_probe.so (linked with host_util.o for the tensor-map helpers)."""
    srcs = sorted(glob.glob(os.path.join(TEST_NATIVE, "*.cu")))
    if not srcs:
        return
    newest = max([hdr_t] + [os.path.getmtime(s) for s in srcs])
    if not force and os.path.exists(TEST_LIB) and os.path.getmtime(TEST_LIB) >= newest:
        return
    cmd = [NVCC, *FLAGS, "-I", CSRC, "-shared", "-o", TEST_LIB, *srcs, os.path.join(OBJ, "host_util.o")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed for tests/native")


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv, ptxas_v="--ptxas" in sys.argv))
