"""Statistics of the folded chain (csrc/chain_fast.cuh) against the op-by-op chain (csrc/chain.cuh) in the same kernel:
random rotations / polynomials / sizes, full-size outputs on noise sources (bilinear: a coordinate that moves by 1/32 px
changes the result).  Prints the number of differing output bytes per case and in total.  Design aid, not a test:
   python scripts/fast_chain_sweep.py [cases] [seed]"""
import sys
sys.path.insert(0, ".")
import numpy as np
import torch
import vr180_convert_b200 as V

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
lib = V._native.lib()
total_px = total_bad = 0
for c in range(cases):
    n = int(rng.choice([1024, 2048, 3072, 4096]))
    hin = win = int(rng.choice([n // 2, n // 2 + 37, n]))
    ang = float(rng.choice([0.0, 0.01, 0.1, 0.5, 1.5]))
    ts = []
    for eye in range(2):
        t = V.EquirectangularEncoder(is_latitude_y=bool(rng.random() < 0.8))
        if ang > 0:
            ax = rng.normal(size=3); ax /= np.linalg.norm(ax)
            a = ang * float(rng.uniform(0.5, 1.0))
            t = t * V.Euclidean3DRotator(V.quaternion(np.cos(a / 2), *(np.sin(a / 2) * ax)))
        if rng.random() < 0.7:
            t = t * V.PolynomialScaler([float(rng.normal(0, 0.01)) if rng.random() < 0.3 else 0.0, float(rng.uniform(0.8, 1.2)),
                                        float(rng.normal(0, 0.03)), float(rng.normal(0, 0.005))])
        ts.append(t * V.FisheyeDecoder("equidistant"))
    radius = float(rng.uniform(0.4, 0.6) * hin)
    g = torch.Generator(device="cuda").manual_seed(c)
    left = torch.randint(0, 256, (1, hin, win, 3), dtype=torch.uint8, device="cuda", generator=g)
    right = torch.randint(0, 256, (1, hin, win, 3), dtype=torch.uint8, device="cuda", generator=g)
    wp = V.SbsWarper((ts[0], ts[1]), size_input=(hin, win), size_output=(n, n), interpolation=1, radius=radius, map_source="analytic")
    lib.vr180_debug_set(1, 256)
    ref = wp(left, right).clone()
    lib.vr180_debug_set(1, 0)
    got = wp(left, right)
    bad = int((got != ref).sum())
    px = 2 * n * n
    total_px += px
    total_bad += bad
    print(f"case {c:3d} n={n} in={hin} angle<={ang} radius={radius:.1f}: {bad} differing bytes of {3 * px}", flush=True)
lib.vr180_debug_set(1, -1)
print(f"TOTAL {total_bad} differing bytes over {total_px / 1e6:.0f} M output pixels ({cases} cases)")
