"""TEST INFRASTRUCTURE: builds tests/emul/_build/libgtsb_sim.so -- the product's CUDA sources
(gt-scaffold_b200/csrc/*.cu) compiled by g++ against tests/emul/cusim/ (a functional host model of
the CUDA runtime and device language: fibers for threads, rendezvous for barriers and warp
collectives).  Three rewrites are applied to a copy of every source, nothing else:

    kernel<<<grid, block, smem, stream>>>(args)   ->  cusim::Launch(grid, block, smem, stream)(kernel)(args)
    cudaLaunchCooperativeKernel((const void *) k, ...) -> cusim::coop_launch(k, ...)
    extern __shared__ <attrs> T name[];            ->  T *name = (T *) cusim::B->dyn_smem;

The library exports the same C ABI as libgtscaffold_b200.so; tests/test_sim.py drives it through
gt-scaffold_b200/api.py by handing it to the ctypes loader explicitly.  Nothing in the package
knows about it: the product has no CPU path.
"""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.normpath(os.path.join(HERE, "..", "..", "gt-scaffold_b200", "csrc"))
SIM = os.path.join(HERE, "cusim")
BUILD = os.path.join(HERE, "_build", "sim")
OUT = os.path.join(HERE, "_build", "libgtsb_sim.so")
NCCL = os.path.join(HERE, "_build", "libnccl_cusim.so")      # the rendezvous stand-in gtsb_dist.cu dlopens
UNITS = ["gtsb_api", "gtsb_build", "gtsb_build2", "gtsb_filter", "gtsb_dist", "gtsb_parse", "gtsb_format", "gtsb_mle"]

_LAUNCH = re.compile(r"([A-Za-z_][A-Za-z0-9_:]*(?:<[^<>;(){}]*>)?)\s*<<<(.*?)>>>\s*\(", re.S)
_COOP = re.compile(r"cudaLaunchCooperativeKernel\(\(const void \*\)\s*")
_EXTERN = re.compile(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([A-Za-z_][A-Za-z0-9_]*)\s+([A-Za-z_][A-Za-z0-9_]*)\[\];")


def rewrite(text):
    def launch(m):
        return "cusim::Launch(%s)(%s)(" % (m.group(2), m.group(1))
    text = _LAUNCH.sub(launch, text)
    text = _COOP.sub("cusim::coop_launch(", text)
    text = _EXTERN.sub(lambda m: "%s *%s = (%s *) cusim::B->dyn_smem;" % (m.group(1), m.group(2), m.group(1)), text)
    assert "<<<" not in text, "a launch the rewrite did not catch"
    return text


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h", ".c")))


def build(force=False, asan=None):
    """asan (default: CUSIM_ASAN=1 in the environment): an AddressSanitizer build, libgtsb_sim_asan.so --
    device memory is host memory here, so a kernel that reads or writes outside an allocation is
    reported with its source line.  Run with LD_PRELOAD=libasan.so and ASAN_OPTIONS=detect_leaks=0."""
    global BUILD, OUT, NCCL
    if asan is None:
        asan = os.environ.get("CUSIM_ASAN") == "1"
    if asan:
        BUILD = os.path.join(HERE, "_build", "sim_asan")
        OUT = os.path.join(HERE, "_build", "libgtsb_sim_asan.so")
        NCCL = os.path.join(HERE, "_build", "libnccl_cusim_asan.so")
    deps = [os.path.join(CSRC, f) for f in sources()] + [os.path.join(SIM, f) for f in os.listdir(SIM)] + [__file__]
    if (not force and os.path.exists(OUT) and os.path.exists(NCCL)
            and min(os.path.getmtime(OUT), os.path.getmtime(NCCL)) >= max(os.path.getmtime(d) for d in deps)):
        return OUT
    os.makedirs(BUILD, exist_ok=True)
    for f in sources():
        src = open(os.path.join(CSRC, f)).read()
        name = f[:-3] + ".cpp" if f.endswith(".cu") else f
        if f == "gtsb_dist.cu":
            assert src.count('dlopen("libnccl.so.2"') == 1
            src = src.replace('dlopen("libnccl.so.2"', 'dlopen("%s"' % NCCL)
        with open(os.path.join(BUILD, name), "w") as o:
            o.write(rewrite(src) if f.endswith((".cu", ".cuh")) else src)
    # the sources name the ABI header relative to csrc ("../../include/..."): point them at the real one
    hdr = os.path.normpath(os.path.join(CSRC, "..", "..", "include", "gtscaffold_b200.h"))
    for f in os.listdir(BUILD):
        if f.endswith((".cpp", ".h", ".cuh")):
            path = os.path.join(BUILD, f)
            text = open(path).read()
            if '"../../include/gtscaffold_b200.h"' in text:
                text = text.replace('"../../include/gtscaffold_b200.h"', '"%s"' % hdr)
                with open(path, "w") as o:
                    o.write(text)
    flags = ["-O1", "-std=c++17", "-fPIC", "-ffp-contract=off", "-w", "-I", SIM, "-I", BUILD]
    if asan:
        flags += ["-g", "-fsanitize=address", "-fno-omit-frame-pointer"]
    objs = []
    import concurrent.futures as cf

    def compile_unit(u):
        obj = os.path.join(BUILD, u + ".o")
        subprocess.run(["g++"] + flags + ["-c", os.path.join(BUILD, u + ".cpp"), "-o", obj], check=True)
        return obj

    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_unit, UNITS))
    subprocess.run(["gcc", "-O2", "-fPIC", "-ffp-contract=off", "-c", os.path.join(BUILD, "gtsb_threshold.c"),
                    "-o", os.path.join(BUILD, "gtsb_threshold.o")], check=True)
    subprocess.run(["g++"] + flags + ["-c", os.path.join(SIM, "cusim.cpp"), "-o", os.path.join(BUILD, "cusim.o")], check=True)
    subprocess.run(["g++"] + flags + ["-shared", os.path.join(SIM, "fake_nccl.cpp"), "-o", NCCL, "-lpthread"], check=True)
    subprocess.run(["g++", "-shared"] + (["-fsanitize=address"] if asan else []) + ["-o", OUT] + objs + [os.path.join(BUILD, "gtsb_threshold.o"),
                                                           os.path.join(BUILD, "cusim.o"), "-lm", "-ldl"], check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
