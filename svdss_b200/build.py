"""In-tree build of libsvdss_b200.so (nvcc, sm_100a only). No JIT cache, no torch extension:
the library is a plain C-ABI shared object so the reference's C++ host code could link it."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsvdss_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
         "-Xcompiler", "-fPIC,-fvisibility=hidden,-O3", "--expt-relaxed-constexpr",
         "-Xptxas", "-v", "-ccbin", "/usr/bin/g++",
         # stream 0 in the library = the calling thread's own default stream: two host threads (one searching batch k + 1,
         # one calling batch k; a BAM reader inflating ahead) overlap on the device instead of queueing on the legacy stream
         "--default-stream", "per-thread"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def host_sources():
    """plain C++ translation units of the library (g++: SIMD intrinsics the nvcc front end need not see)"""
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cpp"))


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + host_sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    deps.append(os.path.join(HERE, "host", "rld.hpp"))
    deps.append(os.path.join(HERE, "..", "include", "svdss_b200.h"))
    deps.append(os.path.abspath(__file__))          # the compiler flags live here
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force=False, verbose=False):
    if not force and not stale():
        return LIB
    objs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    procs = []
    for src in sources():
        obj = os.path.join(HERE, "build", os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [NVCC] + FLAGS + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src in host_sources():
        obj = os.path.join(HERE, "build", os.path.basename(src)[:-4] + ".o")
        objs.append(obj)
        cmd = ["/usr/bin/g++", "-O3", "-std=c++17", "-fPIC", "-fvisibility=hidden", "-pthread", "-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append(out)
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("compiler failed on %s" % src)
    with open(os.path.join(HERE, "build", "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        sys.stderr.write("\n".join(log))
    subprocess.check_call([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-lcudart", "-ccbin", "/usr/bin/g++"])
    return LIB


HOST_BIN = os.path.join(HERE, "SVDSS")


def build_host(force=False):
    """C++14 shell (`SVDSS index | search`) over the C ABI; needs only g++ and zlib."""
    srcs = [os.path.join(HERE, "host", f) for f in ("svdss_main.cpp", "io.hpp", "call.hpp", "clusterer.hpp", "smoother.hpp", "clipper.hpp", "rld.hpp")]
    if (not force and os.path.exists(HOST_BIN)
            and all(os.path.getmtime(HOST_BIN) >= os.path.getmtime(s) for s in srcs + [LIB])):
        return HOST_BIN
    subprocess.check_call(["/usr/bin/g++", "-std=c++14", "-O2", "-Wall", "-fopenmp", "-pthread", "-o", HOST_BIN, srcs[0],
                           "-L" + HERE, "-lsvdss_b200", "-lz", "-Wl,-rpath,$ORIGIN"])
    return HOST_BIN


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose=True))
    print(build_host(force="--force" in sys.argv))
