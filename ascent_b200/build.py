"""Build libvr_b200.so (hand-written CUDA for sm_100a + the C ABI) in-tree with nvcc.

    python -m ascent_b200.build            # incremental
    python -m ascent_b200.build --force

The shared object stays next to this file (git-ignored, shipped to the GPU box by gpurun).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, os.environ.get("VR_LIB_NAME", "libvr_b200.so"))
HOST_OUT = os.path.join(HERE, "libvtkh_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
HOSTCXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# --fmad=false: the sampler and the float fold must round like the reference's scalar x86 code
# (every decision bit-identical to the oracle); explicit __fmaf_rn is used where fusion is wanted.
EXTRA = os.environ.get("VR_NVCC_EXTRA", "").split()
COMMON = EXTRA + ["-O3", "-lineinfo", "-std=c++17", "--fmad=false", "-ccbin", HOSTCXX, "-Xcompiler",
          "-fPIC,-ffp-contract=off,-fvisibility=hidden", "-Xptxas", "-v"]
SOURCES = ["sampler.cu", "stage.cu", "composite.cu", "layers.cu", "comm.cu", "png.cu", "unstructured.cu", "vr_api.cu"]
HEADERS = ["vr_internal.h", "vr_host_math.hpp", "vr_color_table.hpp", "vr_radixk.hpp", "vr_umesh_geom.hpp", "vr_umesh_faces.hpp", os.path.join("..", "..", "include", "vr_b200.h")]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    objs = []
    logs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(CSRC, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _newer(o, [s] + hdrs):
            cmd = [NVCC] + ARCH + COMMON + ["-c", s, "-o", o]
            r = subprocess.run(cmd, capture_output=True, text=True)
            logs.append("$ " + " ".join(cmd) + "\n" + r.stderr)
            if r.returncode != 0:
                sys.stderr.write(logs[-1])
                raise RuntimeError("nvcc failed on " + src)
    if force or _newer(OUT, objs):
        cmd = [NVCC] + ARCH + ["-shared", "-ccbin", HOSTCXX, "-o", OUT] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stderr)
            raise RuntimeError("link failed")
    if logs:
        with open(os.path.join(CSRC, "ptxas.log"), "w") as f:
            f.write("\n".join(logs))
    if verbose:
        print("\n".join(logs))
    build_host(force)
    return OUT


def build_host(force=False):
    """libvtkh_b200.so: the C++ host mirror of vtkh::VolumeRenderer & co. over the C ABI, and the
    C++ test driver that reads like the reference's t_vtk-h_volume_renderer.cpp."""
    hdir = os.path.join(CSRC, "host")
    src = os.path.join(hdir, "vtkh_b200.cpp")
    deps = [src, os.path.join(hdir, "vtkh_b200.hpp"), os.path.join(CSRC, "..", "..", "include", "vr_b200.h")]
    common = [HOSTCXX, "-O2", "-std=c++17", "-fPIC", "-Wall", "-ffp-contract=off"]
    if force or _newer(HOST_OUT, deps + [OUT]):
        cmd = common + ["-shared", "-o", HOST_OUT, src, "-L" + HERE, "-l:" + os.path.basename(OUT),
                        "-Wl,-rpath,$ORIGIN"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stderr)
            raise RuntimeError("host library build failed")
    tsrc = os.path.join(hdir, "t_vtkh_b200_volume_renderer.cpp")
    texe = os.path.join(HERE, "t_vtkh_b200_volume_renderer")
    if os.path.exists(tsrc) and (force or _newer(texe, deps + [tsrc, HOST_OUT])):
        cmd = common + ["-o", texe, tsrc, "-L" + HERE, "-l:libvtkh_b200.so", "-l:" + os.path.basename(OUT),
                        "-Wl,-rpath,$ORIGIN"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stderr)
            raise RuntimeError("host test driver build failed")
    return HOST_OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(OUT)
