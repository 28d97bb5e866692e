"""SwinIR stage-1 timing on a B200 (run through gpurun): ms per image at 512^2 and 1024^2, CUDA events, after warm-up."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import instarevive_b200 as ir  # noqa: E402
from instarevive_b200 import weights  # noqa: E402

dev = torch.device("cuda:0")
net = ir.SwinIR(weights.make_swinir_state_dict(seed=7), device=dev)
for side, B in ((512, 1), (1024, 1), (512, 8)):
    x = torch.rand(B, 3, side, side, device=dev)
    for _ in range(3):
        net(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10):
        net(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"swinir {B}x{side}x{side}: {ms:.2f} ms  ({B * side * side / 1e6 / (ms / 1e3):.1f} MP/s)", flush=True)
