"""Builds libmcedm_b200.so in-tree with nvcc for sm_100a (no torch headers involved).

Each .cu under csrc/ is one translation unit -> csrc/_obj/<name>.o, linked into
mcedm_b200/lib/libmcedm_b200.so.  Objects are rebuilt only when the source (or a header) is newer.
Run as `python -m mcedm_b200.build` or through `__graft_entry__.build()`.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libmcedm_b200.so")
# checker / probe kernels (tests and scripts only): their own library, linked with a private copy of the runtime
CHECK_LIB = os.path.join(LIBDIR, "libmcedm_b200_check.so")
CHECK_SOURCES = ("probe.cu",)
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
    "-I", INCLUDE,
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: the sm_100a kernels cannot be built")
    return nvcc


def _newest_header_mtime() -> float:
    m = 0.0
    for d in (CSRC, INCLUDE):
        for f in os.listdir(d):
            if f.endswith((".cuh", ".h")):
                m = max(m, os.path.getmtime(os.path.join(d, f)))
    return m


def build(verbose: bool = False, force: bool = False, ptxas_info: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    nvcc = _nvcc()
    hdr = _newest_header_mtime()
    sources = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    jobs = []
    objs = []
    for src in sources:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src[:-3] + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr):
            cmd = [nvcc, *NVCC_FLAGS, "-c", s, "-o", o]
            if ptxas_info:
                cmd[1:1] = ["-Xptxas", "-v"]
            jobs.append((src, cmd))

    def run(job):
        name, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        return name, r

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for name, r in ex.map(run, jobs):
            if verbose or r.returncode != 0 or ptxas_info:
                sys.stderr.write(f"[mcedm_b200.build] {name}\n{r.stdout}{r.stderr}\n")
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {name}")
    check_objs = [o for o in objs if os.path.basename(o)[:-2] + ".cu" in CHECK_SOURCES]
    prod_objs = [o for o in objs if o not in check_objs]
    for lib, members in ((LIB, prod_objs), (CHECK_LIB, check_objs + [os.path.join(OBJ, "runtime.o")])):
        relink = bool(jobs) or not os.path.exists(lib) or any(os.path.getmtime(o) > os.path.getmtime(lib) for o in members)
        if relink:
            cmd = [nvcc, "-shared", "-o", lib, *members, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError(f"link of {os.path.basename(lib)} failed")
    return LIB


if __name__ == "__main__":
    path = build(verbose="-v" in sys.argv, force="-f" in sys.argv, ptxas_info="--ptxas" in sys.argv)
    print(path)
