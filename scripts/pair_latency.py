"""Host-array latency of ONE stereo pair through lr_frame (apply_lr's in-memory part): design aid, not a test."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import vr180_convert_b200 as V
from bench import build_transformers
for n in (2048, 4096):
    t = build_transformers(V, "rot_poly", True)
    rng = np.random.default_rng(0)
    l = rng.integers(16, 256, (n, n, 3), dtype=np.uint8); r = rng.integers(16, 256, (n, n, 3), dtype=np.uint8)
    yy, xx = np.ogrid[:n, :n]
    out = (xx - n // 2) ** 2 + (yy - n // 2) ** 2 > (n // 2 - 8) ** 2
    l[out] = 0; r[out] = 0
    lp = V.pinned_empty(l.shape); lp[...] = l; rp = V.pinned_empty(r.shape); rp[...] = r
    for name, a, b in (("pageable", l, r), ("pinned", lp, rp)):
        for interp in (1, 4):
            for radius in (n / 2, "auto"):
                f = lambda: V.lr_frame(t, a, b, size_output=(n, n), interpolation=interp, radius=radius)
                f(); f()
                t0 = time.perf_counter()
                for _ in range(10): f()
                dt = (time.perf_counter() - t0) / 10
                print(f"n={n} {name:8s} interp={interp} radius={radius!s:6s}: {dt*1e3:7.2f} ms per pair  ({2*n*n/1e6/dt:8.0f} Mpix/s)  copy floor {(2*n*n*3*2)/50e9*1e3:.2f} ms")
