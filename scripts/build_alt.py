"""Builds an alternative libmcedm_b200.so with extra -D knobs for A/B timing on one box:
    python scripts/build_alt.py NAME -DMCEDM_FLAT_WG=0 ...   ->  mcedm_b200/lib/alt_NAME.so
then  MCEDM_LIB=mcedm_b200/lib/alt_NAME.so python scripts/prof_fused.py ...   (every .cu is recompiled: ~1 min)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcedm_b200 import build as B  # noqa: E402

name, defs = sys.argv[1], sys.argv[2:]
obj = os.path.join(B.OBJ, "alt_" + name)
os.makedirs(obj, exist_ok=True)
srcs = sorted(f for f in os.listdir(B.CSRC) if f.endswith(".cu"))


def cc(src):
    o = os.path.join(obj, src[:-3] + ".o")
    r = subprocess.run([B._nvcc(), *B.NVCC_FLAGS, *defs, "-c", os.path.join(B.CSRC, src), "-o", o], capture_output=True, text=True)
    if r.returncode:
        sys.exit(r.stdout + r.stderr)
    return o


with ThreadPoolExecutor(8) as ex:
    objs = list(ex.map(cc, srcs))
out = os.path.join(B.LIBDIR, f"alt_{name}.so")
subprocess.run([B._nvcc(), "-shared", "-o", out, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"], check=True)
print(out)
