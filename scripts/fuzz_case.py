"""Re-run one case of tests/test_gpu_fuzz.py with variations and print where it differs (debug aid)."""
import sys
import cv2, numpy as np, torch
sys.path.insert(0, ".")
import vr180_convert_b200 as V
from oracle import chain_np
from tests.test_gpu_fuzz import _draw_chain, BORDERS

seed = int(sys.argv[1])
rng = np.random.default_rng(1000 + seed)
hin, win = int(rng.integers(40, 260)), int(rng.integers(40, 300))
wout = int(rng.choice([16, 32, 64, 96, 100, 128, 136, 160, 208, 256]))
hout = int(rng.choice([8, 16, 24, 32, 40, 64, 72, 96, 104, 128]))
interp = int(rng.choice([0, 1, 1, 2, 4]))
border = BORDERS[int(rng.integers(0, len(BORDERS)))] if rng.random() < 0.5 else cv2.BORDER_CONSTANT
value = tuple(int(v) for v in rng.integers(0, 256, 3)) if rng.random() < 0.3 else (0, 0, 0)
n_frames = int(rng.choice([1, 1, 2, 3, 5, 7, 14, 23]))
per_eye = bool(rng.random() < 0.5)
source = str(rng.choice(["auto", "analytic", "lut", "lut_fixed", "lut_packed", "lut_packed"]))
radius = float(rng.uniform(0.35, 0.8) * min(hin, win))
tl, ops_l = _draw_chain(rng)
tr, ops_r = _draw_chain(rng) if per_eye else (tl, ops_l)
ln = rng.integers(0, 256, (n_frames, hin, win, 3), dtype=np.uint8)
rn = rng.integers(0, 256, (n_frames, hin, win, 3), dtype=np.uint8)
print(dict(hin=hin, win=win, wout=wout, hout=hout, interp=interp, border=border, value=value, n_frames=n_frames, per_eye=per_eye, source=source, radius=radius))
print(ops_l, ops_r)
ml = chain_np.get_map(ops_l, radius=radius, size_input=(hin, win), size_output=(wout, hout))
mr = chain_np.get_map(ops_r, radius=radius, size_input=(hin, win), size_output=(wout, hout)) if per_eye else ml
left, right = torch.from_numpy(ln).cuda(), torch.from_numpy(rn).cuda()
for src in ("analytic", "lut", "lut_fixed", "lut_packed"):
    wp = V.SbsWarper((tl, tr) if per_eye else tl, size_input=(hin, win), size_output=(wout, hout), interpolation=interp,
                     radius=radius, map_source=src, boarder_mode=border, boarder_value=value)
    got = wp(left, right).cpu().numpy()
    gm = wp.maps().cpu().numpy() if src != "analytic" else None
    for f in range(n_frames):
        want = np.concatenate([cv2.remap(ln[f], ml[0], ml[1], interpolation=interp, borderMode=border, borderValue=value),
                               cv2.remap(rn[f], mr[0], mr[1], interpolation=interp, borderMode=border, borderValue=value)], axis=1)
        bad = np.argwhere((got[f] != want).any(axis=2))
        if len(bad):
            print(src, "frame", f, "bad pixels", bad[:8].tolist(), "got", got[f][tuple(bad[0])], "want", want[tuple(bad[0])])
            y, x = bad[0]
            m = ml if x < wout else mr
            xx = x % wout
            print("   map at first bad:", m[0][y, xx], m[1][y, xx], "x*32", m[0][y, xx] * 32, m[1][y, xx] * 32)
            if gm is not None:
                k = 0 if (x < wout or not per_eye) else 1
                print("   device map:", gm[k, 0, y, xx], gm[k, 1, y, xx])
    print(src, "done")
