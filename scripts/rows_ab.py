"""conv_rows_fused A/B under MCEDM_DBG switches: python scripts/rows_ab.py [B] [dbg values...]"""
import os, subprocess, sys
B = sys.argv[1] if len(sys.argv) > 1 else "256"
here = os.path.dirname(os.path.abspath(__file__))
for dbg in (sys.argv[2:] or ["8", "0"]):
    for xf, res in (("1", "0"), ("1", "1")):
        subprocess.run([sys.executable, os.path.join(here, "rows_ablate.py"), B, xf, res], env=dict(os.environ, MCEDM_DBG=dbg))
    env = dict(os.environ, MCEDM_DBG=dbg)
    subprocess.run([sys.executable, os.path.join(here, "rows_ctr.py"), B, "x"], env=env)
