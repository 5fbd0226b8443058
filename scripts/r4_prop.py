"""Propagation timing of the session-4 kernels (cp.async-prefetching rows passes, twiddles in shared memory,
late PDL wait) against the previous library and with each change switched off."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUN = os.path.join(ROOT, "scripts", "r2_prop.py")
prev = os.path.join(ROOT, "build", "exp", "lib_prev.so")
variants = [{}, {"BHOLO_FFT_B": "3"}, {"BHOLO_FFT_C": "2"}, {"BHOLO_FFT_B": "3", "BHOLO_FFT_C": "2"},
            {"BHOLO_FFT_NO_LATE_WAIT": "1"}]
if os.path.exists(prev):
    variants.append({"BHOLO_LIB": prev, "BHOLO_FFT_A": "2", "BHOLO_FFT_C": "2"})
for N in (1024, 896):
    for extra in variants:
        env = dict(os.environ, **extra)
        r = subprocess.run([sys.executable, RUN, "run", str(N)], env=env, capture_output=True, text=True)
        print(extra, r.stdout.strip() or r.stderr[-1500:], flush=True)
