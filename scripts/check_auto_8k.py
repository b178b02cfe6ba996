"""GPU check: per-frame auto radius at 8K == the same radius baked into the chain (tiled DYN vs non-DYN kernels)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
import vr180_convert_b200 as V  # noqa: E402

n = 4096
dev = torch.device("cuda", 0)
left = bench.synth_frames_torch(torch, 3, n, 1, dev)
right = bench.synth_frames_torch(torch, 3, n, 2, dev)
t = V.EquirectangularEncoder() * V.FisheyeDecoder("equidistant")
for interp in (1, 2):
    wa = V.SbsWarper(t, size_input=(n, n), size_output=(n, n), interpolation=interp, radius="auto")
    rad, trans = wa.radius_per_frame(left, right)
    print("radii", rad.cpu().tolist(), trans.cpu().tolist()[0])
    oa = wa(left, right)
    wf = V.SbsWarper(t, size_input=(n, n), size_output=(n, n), interpolation=interp, radius=float(rad[0].item()))
    of = wf(left, right)
    print("interp", interp, "equal", torch.equal(oa, of), "nonzero frac auto", float((oa != 0).float().mean()),
          "fixed", float((of != 0).float().mean()), "mismatch px", int((oa != of).any(dim=-1).sum()))
