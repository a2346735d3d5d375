"""Build recipe of libharcgpu.so and the two drop-in executables (in-tree, sm_100a only)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libharcgpu.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CUFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "--expt-relaxed-constexpr",
           "-ccbin", "/usr/bin/g++"] + os.environ.get("HARC_CUFLAGS", "").split()
CU = ["capi.cu", "stage1.cu", "walk.cu", "stage2.cu", "scan.cu", "ingest.cu", "sort.cu", "job.cu"]
HDR = ["common.cuh", "ctx.h", os.path.join("..", "..", "include", "harcgpu.h")]
EXES = {"reorder.out": "reorder_main.cpp", "encoder.out": "encoder_main.cpp"}


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(verbose=False, force=False):
    objs = []
    hdrs = [os.path.join(CSRC, h) for h in HDR]
    procs = []
    for cu in CU:
        src = os.path.join(CSRC, cu)
        obj = os.path.join(CSRC, cu[:-3] + ".o")
        objs.append(obj)
        if force or _newer(obj, [src] + hdrs):
            cmd = [NVCC] + ARCH + CUFLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for cmd, p in procs:
        out = p.communicate()[0].decode()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    if force or _newer(LIB, objs):
        cmd = [NVCC] + ARCH + ["-shared", "-ccbin", "/usr/bin/g++", "-o", LIB] + objs + ["-lcudart"]
        subprocess.check_call(cmd)
    for exe, src in EXES.items():
        s = os.path.join(CSRC, src)
        t = os.path.join(HERE, exe)
        if os.path.exists(s) and (force or _newer(t, [s, LIB])):
            subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", s, "-o", t, "-I", os.path.join(HERE, "..", "include"),
                                   "-L", HERE, "-lharcgpu", "-Wl,-rpath,$ORIGIN"])
    return LIB


if __name__ == "__main__":
    build(verbose="-v" in sys.argv, force="-f" in sys.argv)
    print(LIB)
