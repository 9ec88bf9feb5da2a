"""Times the fused pyramid kernel (2 images 752x480 per launch, and 32 per launch; device-resident frames) for the library named by
HV_LIB_PATH and checks one pyramid against the C oracle (bit-exact). HV_PYR_V2=1 selects hv_pyr_fused2_kernel:
    python tests/tools/pyr_time.py; HV_PYR_V2=1 python tests/tools/pyr_time.py"""
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
# 32 images per launch (the kernels_batched line of bench.py): 16 stereo pairs, every image into its own pyramid
big = [hv.pyramid(W, H) for _ in range(32)]
imgs = [frames[i % 16, i // 16] for i in range(32)]
for i in range(3):
    hv.build_pyramids(big, imgs, device=True)
hv.sync()
with torch.cuda.stream(st):
    s.record(st)
    for i in range(50):
        hv.build_pyramids(big, imgs, device=True)
    e.record(st)
e.synchronize()
us = s.elapsed_time(e) * 1e3 / 50
print(f"{'hv_pyr_fused2_kernel' if os.environ.get('HV_PYR_V2') else 'hv_pyr_fused_kernel'}: 32 images per launch {us:.2f} us = {32 * 2397000 / us * 1e-3:.0f} GB/s algorithmic")
