"""Times hv_pyr_fused_kernel (2 images 752x480, device-resident frames) for the library named by HV_LIB_PATH and checks
one pyramid against the C oracle (bit-exact)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from hybvio_b200 import capi, synth
from oracle import lk_oracle
hv = capi.Context(0)
W, H = 752, 480
frames = synth.stereo_frames_torch(0, 16, W, H, seed=42, device=torch.device("cuda", 0))
pyr = [hv.pyramid(W, H) for _ in range(2)]
for i in range(5):
    hv.build_pyramids(pyr, [frames[i, 0], frames[i, 1]], device=True)
hv.sync()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
st = torch.cuda.ExternalStream(hv.stream)
with torch.cuda.stream(st):
    s.record(st)
    for i in range(200):
        hv.build_pyramids(pyr, [frames[i % 16, 0], frames[i % 16, 1]], device=True)
    e.record(st)
e.synchronize()
orc = lk_oracle.OracleLK()
img = frames[7, 0].cpu().numpy()
hv.build_pyramids(pyr, [frames[7, 0], frames[7, 1]], device=True)
o = orc.pyramid(img)
ok = True
for lv in range(o.levels):
    g, d = pyr[0].download(lv)
    og, od = o.download(lv, padded=False)
    ok = ok and np.array_equal(g, og) and np.array_equal(d, od)
print(f"{os.path.basename(capi.LIB_PATH)}: pyramid pair {s.elapsed_time(e) * 1e3 / 200:.2f} us per launch, bit-exact vs oracle: {ok}")
