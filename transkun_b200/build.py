"""Builds transkun_b200/csrc/libtranskun_b200.so in-tree with nvcc for sm_100a.

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
nvcc cross-compiles without a GPU, so this also is the CPU-side "does it build" check.
"""
from __future__ import annotations

import glob
import os
import subprocess

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB = os.path.join(CSRC, "libtranskun_b200.so")
INCLUDE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")

# flags: -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC; linked -shared with
# -cudart shared (the same libcudart the host process, PyTorch, already loaded)


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h"))
    return any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps)


def build(force: bool = False, verbose: bool = False, defines=(), out: str = LIB) -> str:
    """Compiles every .cu to an object file in parallel (the two sweep kernels dominate), then links."""
    if out == LIB:
        build_pylists(force)
    if not force and out == LIB and not needs_build():
        return LIB
    import concurrent.futures
    import tempfile
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    common = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
              *[f"-D{d}" for d in defines], "-I", INCLUDE]
    if verbose:
        common[1:1] = ["-Xptxas", "-v"]
    log = []
    with tempfile.TemporaryDirectory(prefix="tkb_obj_") as tmp:
        def compile_one(src):
            obj = os.path.join(tmp, os.path.basename(src)[:-3] + ".o")
            res = subprocess.run([*common, "-c", src, "-o", obj], capture_output=True, text=True)
            if res.returncode != 0:
                raise RuntimeError("nvcc failed:\n" + " ".join([*common, "-c", src]) + "\n" + res.stdout + res.stderr)
            return obj, res.stderr
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
            objs = []
            for obj, err in ex.map(compile_one, sources()):
                objs.append(obj)
                log.append(err)
        link = [nvcc, "-shared", "-cudart", "shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out, *objs,
                "-lcufft"]
        res = subprocess.run(link, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n" + " ".join(link) + "\n" + res.stdout + res.stderr)
    if verbose:
        print("".join(log))
    return out


PYLISTS = os.path.join(CSRC, "_tkb_pylists.so")


def build_pylists(force: bool = False) -> str:
    """Host-side CPython helper (csrc/pylists.c): packed decode result -> the reference's list-of-lists-of-tuples."""
    import sysconfig
    src = os.path.join(CSRC, "pylists.c")
    if not force and os.path.exists(PYLISTS) and os.path.getmtime(PYLISTS) >= os.path.getmtime(src):
        return PYLISTS
    cmd = [os.environ.get("CC", "gcc"), "-O2", "-shared", "-fPIC", "-I", sysconfig.get_paths()["include"], src, "-o", PYLISTS]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("gcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    return PYLISTS


def build_timeline() -> str:
    """Diagnostics build: same sources with -DTKB_TIMELINE (per-block globaltimer stamps)."""
    return build(force=True, defines=("TKB_TIMELINE",), out=os.path.join(CSRC, "libtranskun_b200_timeline.so"))


if __name__ == "__main__":
    print(build(force=True, verbose=True))
