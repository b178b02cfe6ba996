import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import vr180_convert_b200 as V
n = 4096
t = V.EquirectangularEncoder() * V.FisheyeDecoder("equidistant")
g = torch.Generator(device='cuda'); g.manual_seed(0)
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 4
left = torch.randint(16, 256, (pairs, n, n, 3), dtype=torch.uint8, device='cuda', generator=g)
right = torch.randint(16, 256, (pairs, n, n, 3), dtype=torch.uint8, device='cuda', generator=g)
for interp, name in ((0, 'nearest'), (1, 'linear'), (2, 'cubic'), (4, 'lanczos4')):
    wp = V.SbsWarper(t, size_input=(n, n), size_output=(n, n), interpolation=interp, radius=n / 2)
    out = wp(left, right); torch.cuda.synchronize()  # also allocates the output once: the timed calls reuse it
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): wp(left, right, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name:9s} {ms:8.3f} ms per {pairs} pairs  {pairs * n * 2 * n / ms / 1e3:9.1f} Mpix/s")
