"""Propagation timing at 1024^2 x 24 and 896^2 x 24 (GPU): register-resident passes (bh_fft2.cuh, default),
the same launched colour group by colour group (BHOLO_FFT_GROUPED=1) and the shared-memory passes of round 1
(BHOLO_FFT_V1=1); plus the exhaustive sweep (complex-input passes) and the PSNR each variant reports."""
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if len(sys.argv) > 1 and sys.argv[1] == "run":
    import numpy as np
    import torch
    import binary_hologram_reinforcement_learning_b200 as bh
    N = int(sys.argv[2]); F = 24
    eng = bh.HoloEngine(N, F, bh.WL_RGB, n_env=2)
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)
    for e in range(2):
        pre, tgt = bh.synthetic_problem(N, F, 3, e)
        eng.set_target(e, tgt); eng.load_state(e, (pre >= 0.5).astype(np.int8))
    psnr = eng.metrics(0)[0]
    ms = min(eng.time_propagate(0, 20) for _ in range(3))
    passes = eng.time_propagate_passes(0, 5)
    d_map = torch.empty(F * N * N, dtype=torch.float64, device="cuda")
    eng.sweep_all_device(0, d_map.data_ptr()); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        eng.sweep_all_device(0, d_map.data_ptr())
    torch.cuda.synchronize()
    sweep_ms = (time.perf_counter() - t0) / 3 * 1e3
    bytes_survey = F * (N * N + 32.0 * N * N) + 8.0 * 3 * N * N
    print(json.dumps({"N": N, "v1": bool(os.environ.get("BHOLO_FFT_V1")), "mode": os.environ.get("BHOLO_FFT_MODE", "1"),
                      "ABC": os.environ.get("BHOLO_FFT_A", "2") + os.environ.get("BHOLO_FFT_B", "3") + os.environ.get("BHOLO_FFT_C", "2"), "W": os.environ.get("BHOLO_FFT_W", "8"),
                      "rev": bool(os.environ.get("BHOLO_FFT_REV")), "lib": os.path.basename(os.environ.get("BHOLO_LIB", "default")),
                      "psnr": psnr, "propagate_ms": round(ms, 4), "passes_ms": [round(x, 4) for x in passes],
                      "gbs_survey_model": round(bytes_survey / ms / 1e6), "sweep_all_ms": round(sweep_ms, 3),
                      "sweep_checksum": float(d_map[::4097].sum().item())}))
else:
    M0, M2 = {"BHOLO_FFT_MODE": "0"}, {"BHOLO_FFT_MODE": "2"}
    SC = {"BHOLO_LIB": os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "build", "exp", "lib_scalar.so")}
    if not os.path.exists(SC["BHOLO_LIB"]):
        SC = None
    variants = [{}, M2, M0, {"BHOLO_FFT_B": "2"}, {"BHOLO_FFT_C": "3"}] + ([SC, dict(SC, **M0)] if SC else [])
    for N in (1024, 896):
        for extra in (variants if N == 1024 else [{}, M2] + ([SC] if SC else [])):
            env = dict(os.environ, **extra)
            r = subprocess.run([sys.executable, __file__, "run", str(N)], env=env, capture_output=True, text=True)
            print(r.stdout.strip() or r.stderr[-1500:], flush=True)
