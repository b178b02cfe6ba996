"""Design-time simulation (CPU, NumPy) of the bilinear row-patch tap loads of csrc/tiled.cu on a real map:
shared-memory wavefronts per warp step (32 pixels) for the shipped layout and for candidate variants.
Not product code, not a test.

    python scripts/sim_rowpatch.py /tmp/map2048.npz
"""
from __future__ import annotations

import sys

import numpy as np


def degree(words: np.ndarray, active: np.ndarray | None = None) -> np.ndarray:
    """words: (N, 32) word addresses of N warp-wide LDS.32 -> wavefronts each (max distinct words per bank)."""
    N = words.shape[0]
    w = words if active is None else np.where(active, words, -1)
    a = np.sort(w, axis=1)
    first = np.ones_like(a, dtype=bool)
    first[:, 1:] = a[:, 1:] != a[:, :-1]
    first &= a >= 0
    cnt = np.zeros((N, 32), dtype=np.int32)
    rows = np.broadcast_to(np.arange(N)[:, None], a.shape)
    np.add.at(cnt, (rows[first], (a % 32)[first]), 1)
    d = cnt.max(axis=1)
    if active is not None:
        d = np.where(active.any(axis=1), np.maximum(d, 1), 0)
    return d


def main():
    z = np.load(sys.argv[1])
    xm, ym = z["xm"], z["ym"]
    n = xm.shape[0]
    sx = np.rint(xm.astype(np.float64) * 32).astype(np.int64) >> 5
    sy = np.rint(ym.astype(np.float64) * 32).astype(np.int64) >> 5
    T = 32
    ty = tx = n // T
    bx = sx.reshape(ty, T, tx, T).transpose(0, 2, 1, 3).reshape(-1, T, T)  # (tile, row, col)
    by = sy.reshape(ty, T, tx, T).transpose(0, 2, 1, 3).reshape(-1, T, T)
    rng = np.random.default_rng(0)
    sel = rng.choice(bx.shape[0], size=min(3000, bx.shape[0]), replace=False)
    bx, by = bx[sel], by[sel]
    ok = (bx.min(axis=(1, 2)) > 0) & (by.min(axis=(1, 2)) > 0) & (bx.max(axis=(1, 2)) < n - 2) & (by.max(axis=(1, 2)) < n - 2)
    bx, by = bx[ok], by[ok]
    nt = bx.shape[0]
    mnx, mxx = bx.min(axis=(1, 2)), bx.max(axis=(1, 2))
    mny, mxy = by.min(axis=(1, 2)), by.max(axis=(1, 2))
    bx0 = (3 * mnx) & ~15
    bx1 = (3 * (mxx + 2) + 15) & ~15
    wbytes = bx1 - bx0
    nrows = mxy + 2 - mny
    box_rows = 32 + np.maximum(0, (nrows - 32 + 3) // 4) * 4
    print(f"{nt} tiles; wbytes mean {wbytes.mean():.1f}; nrows mean {nrows.mean():.1f}; box rows mean {box_rows.mean():.1f}")
    col = 3 * bx - bx0[:, None, None]  # byte column of tap 00
    row = by - mny[:, None, None]

    def loads(pitch):  # pitch: (nt,) -> wavefronts (nt, 32 steps, 6 loads)
        byte0 = row * pitch[:, None, None] + col
        w0 = (byte0 >> 2).reshape(-1, 32)
        third = ((byte0 & 3) == 3).reshape(-1, 32)
        pw = np.repeat(pitch // 4, 32)[:, None]
        out = []
        for r in (0, 1):
            for j in (0, 1):
                out.append(degree(w0 + r * pw + j))
            out.append(degree(w0 + r * pw + 2, third))
        return np.stack(out, axis=1).reshape(nt, 32, 6)

    base_pitch = np.maximum(wbytes, 96)
    # shipped choice: 4 candidates, cost = first tap load of step k = 0 of every warp x 24 loads + box_rows * pitch / 64
    cand_costs, cand_wf = [], []
    for e in range(8):
        p = base_pitch + 16 * e
        wf = loads(p)
        cand_wf.append(wf)
        first = wf[:, ::4, 0].sum(axis=1)  # rows 0, 4, 8, ... (step k = 0 of the 8 warps), load 0
        cand_costs.append(first * 24 + box_rows * p // 64)
    cand_costs = np.stack(cand_costs)  # (8, nt)
    cand_wf = np.stack(cand_wf)        # (8, nt, 32, 6)
    for ncand in (1, 4, 8):
        pick = cand_costs[:ncand].argmin(axis=0)
        wf = cand_wf[pick, np.arange(nt)]
        tma = (box_rows * (base_pitch + 16 * pick) / 64 / 32).mean()
        print(f"cands {ncand}: tap wavefronts/step {wf.sum(axis=2).mean():.3f} (4 full loads {wf[:, :, [0, 1, 3, 4]].sum(axis=2).mean():.3f}, "
              f"pred {wf[:, :, [2, 5]].sum(axis=2).mean():.3f}); TMA-in wavefronts/step {tma:.3f}; extra pitch mean {16 * pick.mean():.1f}")
    # oracle choice (true total cost incl. TMA)
    true_cost = cand_wf.sum(axis=(2, 3)) + (box_rows[None] * (base_pitch[None] + 16 * np.arange(8)[:, None]) / 64)
    for ncand in (4, 8):
        pick = true_cost[:ncand].argmin(axis=0)
        wf = cand_wf[pick, np.arange(nt)]
        tma = (box_rows * (base_pitch + 16 * pick) / 64 / 32).mean()
        print(f"oracle pick of {ncand}: tap {wf.sum(axis=2).mean():.3f}; TMA-in {tma:.3f}")
    # how many source rows does a step touch?
    rows_per_step = np.array([[len(np.unique(row[t, r])) for r in range(32)] for t in range(min(nt, 500))])
    print("source rows per step: mean", rows_per_step.mean(), "hist", np.bincount(rows_per_step.ravel())[:8])


if __name__ == "__main__":
    main()
