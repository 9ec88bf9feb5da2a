"""Times the LK kernel (temporal with initial flow, stereo) on the BASELINE shape."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from hybvio_b200 import capi, synth
s = torch.cuda.Stream(); hv = capi.Context(0, stream=s.cuda_stream)
W, H, N = 752, 480, 150
fr = synth.stereo_frames_torch(0, 2, W, H, device="cuda")
pyr = [hv.pyramid(W, H) for _ in range(4)]
with torch.cuda.stream(s):
    hv.build_pyramids(pyr[0:2], [fr[0, 0], fr[0, 1]], device=True); hv.build_pyramids(pyr[2:4], [fr[1, 0], fr[1, 1]], device=True)
    pts = torch.from_numpy(synth.interior_points(N)).cuda()
    fx, fy = synth.true_flow(0, 1)
    init = (pts + torch.tensor([fx, fy], device="cuda", dtype=torch.float32) + 0.7).contiguous()
    nxt = init.clone(); nxt2 = torch.zeros_like(pts)
    st = torch.zeros(N, dtype=torch.uint8, device="cuda"); ts = torch.zeros(N, dtype=torch.int32, device="cuda")
    for name, fn in (("temporal(init)", lambda: hv.lk_track_device(pyr[0], pyr[2], pts, nxt, st, ts, N, True)),
                     ("stereo", lambda: hv.lk_track_device(pyr[2], pyr[3], pts, nxt2, st, ts, N, False))):
        for _ in range(5): fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s)
        for _ in range(50): fn()
        b.record(s); b.synchronize()
        print(name, "%.1f us" % (a.elapsed_time(b) * 1e3 / 50), "tracked", int(st.sum()))
