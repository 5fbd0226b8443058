"""k_eval with the first units of every CTA's stream requested into L2 before the PDL wait: BHOLO_EVAL_L2PF = 0, 1, 2, 4, 8."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for pf in (3, 4, 5, 6, 0, 4, 3):
    env = dict(os.environ, BHOLO_EVAL_L2PF=str(pf))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "r2_tune.py"), "run"], env=env, capture_output=True, text=True)
    print("L2PF", pf, r.stdout.strip() or r.stderr[-800:], flush=True)
