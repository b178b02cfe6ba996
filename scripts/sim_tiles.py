"""Design-time simulation (CPU, NumPy) of the tiled warp on the real 8K rotated map: per-tile source rectangles and
shared-memory bank conflicts of the tap loads for candidate staging layouts.  Not product code, not a test.

    python scripts/sim_tiles.py [n] [--base]

Layouts compared (a warp samples an 8 x 4 output patch per load instruction):
  rgb   3-byte pixels as TMA lands them, row pitch 160 B: loads r0[0], r0[1], r1[0], r1[1] (+ predicated third)
  rgbx  4-byte pixels (converted in shared memory), row pitch P words: loads p00, p01, p10, p11
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bench import oracle_ops  # noqa: E402
from oracle import chain_np  # noqa: E402


def conflict_degree(addr: np.ndarray, active: np.ndarray | None = None) -> np.ndarray:
    """addr: (P, 32) word addresses of one warp-wide LDS.32 -> wavefronts per instruction (max distinct words in a bank)."""
    P = addr.shape[0]
    a = np.sort(np.where(active, addr, -1) if active is not None else addr, axis=1)
    first = np.ones_like(a, dtype=bool)
    first[:, 1:] = a[:, 1:] != a[:, :-1]
    first &= a >= 0
    cnt = np.zeros((P, 32), dtype=np.int32)
    rows = np.repeat(np.arange(P), 32).reshape(P, 32)
    np.add.at(cnt, (rows[first], (a % 32)[first]), 1)
    return cnt.max(axis=1)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 4096
    kind = "base" if "--base" in sys.argv else "rot_poly"
    xm, ym = chain_np.get_map(oracle_ops(kind, False), radius=n / 2, size_input=(n, n), size_output=(n, n))
    sx = np.rint(xm.astype(np.float64) * 32).astype(np.int64) >> 5
    sy = np.rint(ym.astype(np.float64) * 32).astype(np.int64) >> 5
    T = 32
    ty, tx = n // T, n // T
    bx = sx.reshape(ty, T, tx, T)
    by = sy.reshape(ty, T, tx, T)
    mnx, mxx = bx.min(axis=(1, 3)), bx.max(axis=(1, 3))
    mny, mxy = by.min(axis=(1, 3)), by.max(axis=(1, 3))
    wpx = mxx - mnx + 2
    nrows = mxy - mny + 2
    wbytes = ((3 * (mxx + 2) + 15) & ~15) - ((3 * mnx) & ~15)
    print(f"tiles {ty}x{tx}: wpx mean {wpx.mean():.1f} max {wpx.max()}  nrows mean {nrows.mean():.1f} max {nrows.max()}"
          f"  wbytes mean {wbytes.mean():.1f} max {wbytes.max()}")
    for q in (50, 90, 99, 99.9):
        print(f"  p{q}: wpx {np.percentile(wpx, q):.0f} nrows {np.percentile(nrows, q):.0f} wbytes {np.percentile(wbytes, q):.0f}"
              f" conv bytes {np.percentile(wpx * nrows * 4, q):.0f}")
    raw = np.where(wbytes <= 160, 160, 224) * (32 + 8 * np.maximum(0, (nrows - 32 + 7) // 8))
    conv = wpx * nrows * 4
    print(f"raw stage bytes mean {raw.mean():.0f} max {raw.max()}; conv bytes mean {conv.mean():.0f} max {conv.max()}")
    for budget in (40960, 45056, 49152):
        for s_raw, c in ((2, 2), (3, 2), (3, 3), (4, 3)):
            print(f"  budget {budget}: S_raw={s_raw} C={c} fits {np.mean(s_raw * raw + c * conv <= budget) * 100:.2f}% of tiles")

    # ---- bank conflicts: patches of 8 x 4 output pixels --------------------------------------------------
    # patch p = (tile, band 0..7, k 0..3): rows 4 band .. 4 band + 3, columns 8 k .. 8 k + 7; lane = 8 r + c
    px = bx.reshape(ty, 8, 4, tx, 4, 8).transpose(0, 3, 1, 4, 2, 5).reshape(ty * tx, 32, 32)  # (tile, patch, lane)
    py = by.reshape(ty, 8, 4, tx, 4, 8).transpose(0, 3, 1, 4, 2, 5).reshape(ty * tx, 32, 32)
    ox, oy = mnx.reshape(-1, 1, 1), mny.reshape(-1, 1, 1)
    rel_x, rel_y = px - ox, py - oy
    rng = np.random.default_rng(0)
    sel = rng.choice(ty * tx, size=min(4096, ty * tx), replace=False)
    rel_x, rel_y = rel_x[sel].reshape(-1, 32), rel_y[sel].reshape(-1, 32)
    wsel = np.repeat(wpx.reshape(-1)[sel], 32)
    bx0 = ((3 * mnx) & ~15).reshape(-1)[sel]
    off3 = np.repeat((3 * mnx.reshape(-1)[sel] - bx0), 32).reshape(-1, 1)

    # current layout: rgb bytes, pitch 160
    for pitch in (160, 224):
        byte0 = rel_y * pitch + 3 * rel_x + off3
        w0 = byte0 >> 2
        third = (byte0 & 3) == 3
        tot = 0.0
        for row in (0, 1):
            for j in (0, 1):
                tot += conflict_degree(w0 + row * (pitch // 4) + j).mean()
            tot += conflict_degree(w0 + row * (pitch // 4) + 2, third).mean()
        print(f"rgb pitch {pitch}: wavefronts per pixel-step (6 LDS) {tot:.2f}")

    # rgbx layouts: pitch in words
    def rgbx(pitch_words):
        w0 = rel_y * pitch_words + rel_x
        return np.stack([conflict_degree(w0 + r * pitch_words + j) for r in (0, 1) for j in (0, 1)]).sum(axis=0)

    best = None
    for label, pw in (("wpx", wsel), ("wpx|1", wsel | 1), ("wpx odd>=", wsel + (1 - wsel % 2)),
                      *[(f"{c}", np.full_like(wsel, c)) for c in (33, 37, 40, 41, 44, 45, 47, 49, 52, 56, 57, 64, 65, 72, 73)]):
        d = rgbx(pw.reshape(-1, 1))
        print(f"rgbx pitch {label}: wavefronts per pixel-step (4 LDS) {d.mean():.3f}")
    cands = [wsel.reshape(-1, 1) + e for e in range(0, 9)]
    ds = np.stack([rgbx(c) for c in cands])  # (cand, patches)
    per_tile = ds.reshape(len(cands), -1, 32).mean(axis=2)  # (cand, tile)
    print(f"rgbx best-of wpx+0..8 per tile: {per_tile.min(axis=0).mean():.3f}  (extra words mean "
          f"{per_tile.argmin(axis=0).mean():.2f})")
    # LDS.64 on pair cells (8 B per source pixel: [p(x), p(x+1)]): two half-warp phases
    def pair64(pitch_cells):
        tot = 0
        c0 = rel_y * pitch_cells + rel_x
        for r in (0, 1):
            a = (c0 + r * pitch_cells) * 2
            for half in (slice(0, 16), slice(16, 32)):
                both = np.concatenate([a[:, half], a[:, half] + 1], axis=1)
                pad = np.full((a.shape[0], 32 - both.shape[1]), -1) if both.shape[1] < 32 else None
                arr = both if pad is None else np.concatenate([both, pad], axis=1)
                tot = tot + conflict_degree(arr, arr >= 0)
        return tot
    for e in (0, 1):
        print(f"pair cells LDS.64, pitch wpx+{e}: wavefronts per pixel-step {pair64(wsel.reshape(-1, 1) + e).mean():.3f}")


if __name__ == "__main__":
    main()
