"""Launches the corner-detection kernel on a BASELINE-config frame (for ncu captures) and prints its CUDA-event time."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from hybvio_b200 import capi, synth
hv = capi.Context(0)
W, H = 752, 480
fr = synth.stereo_frames_torch(0, 2, W, H, device="cuda")
p = hv.pyramid(W, H)
hv.build_pyramids([p], [fr[0, 0]], device=True)
cx, cy = p.gftt_cells(32)
d_kp = torch.zeros((cx * cy, 3), dtype=torch.float32, device="cuda")
for _ in range(3):
    p.gftt_detect_device(d_kp.data_ptr())
hv.sync()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
st = torch.cuda.ExternalStream(hv.stream) if hasattr(hv, "stream") else torch.cuda.current_stream()
torch.cuda.synchronize()
with torch.cuda.stream(st):
    s.record(st)
    for _ in range(50):
        p.gftt_detect_device(d_kp.data_ptr())
    e.record(st)
e.synchronize()
us = s.elapsed_time(e) * 1e3 / 50
print(f"hv_gftt_kernel {W}x{H}, {cx * cy} cells: {us:.2f} us per launch, {W * H / us * 1e-3:.1f} GB/s of algorithmic input bytes; kp[0]={d_kp[0].tolist()}")
