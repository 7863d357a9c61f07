#!/usr/bin/env python
"""FFMA issue rate vs warps per SM (register-only kernel, 16 independent chains per thread)."""
import ctypes as C
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "neural-implicit-queries_b200"))
import _niq
ctx = _niq.Context(0)
L = C.CDLL(_niq.LIB_PATH)
for bps, thr in ((1, 128), (1, 256), (2, 192), (2, 256), (3, 256), (4, 256), (8, 256)):
    out = C.c_float()
    rc = L.niq_probe_ffma(ctx.handle, C.c_int(bps), C.c_int(thr), C.byref(out))
    print(f"{bps} blocks/SM x {thr} threads = {bps * thr // 32:2d} warps/SM ({bps * thr // 128} per scheduler): {out.value:6.2f} TFLOP/s", flush=True)
