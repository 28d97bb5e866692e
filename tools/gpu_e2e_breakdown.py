#!/usr/bin/env python
"""Where the end-to-end time of process() goes at 1024^2: the stages of process() (pipeline.py) run one by one with a
device synchronise after each (wall clock per stage), next to the unsynchronised whole call. usage:
python tools/gpu_e2e_breakdown.py [size]"""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import instarevive_b200 as ir
from instarevive_b200 import pipeline, weights

side = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda:0")
net = ir.ControlPixArtMSHalf(ir.PixArtMS(depth=28, input_size=64, micro_condition=True, init_weights=False), 13).eval()
net.load_state_dict(weights.make_dit_state_dict(depth=28, copy_blocks=13, seed=1), strict=True)
net = net.to(dev)
net.pack()
vae = ir.AutoencoderKL(weights.make_vae_state_dict(dec_seed=2, enc_seed=5), device=dev)
sched = ir.DDPMSchedulerLite()
_, _, y, mask, _ = weights.make_inputs(1, 8, 8, seed=9, lens=(77,))
y, mask = y.to(dev), mask.to(dev)
img = weights.synthetic_degraded_image(side, side, seed=0)
host = torch.from_numpy(np.stack([img])).pin_memory()
host_list = [host[0].numpy()]


def whole():
    return ir.process(net, host_list, strength=1, color_fix_type="wavelet", disable_preprocess_model=True, tiled=False,
                      tile_size=512, tile_stride=448, vae=vae, y=y, y_mask=mask, scheduler=sched, use_control=True)


for _ in range(4):
    whole()
torch.cuda.synchronize()
t0 = time.perf_counter()
N = 10
for _ in range(N):
    whole()
torch.cuda.synchronize()
print(f"process() whole call: {(time.perf_counter() - t0) / N * 1e3:.3f} ms")


def stage(name, fn, acc):
    torch.cuda.synchronize()
    t = time.perf_counter()
    r = fn()
    t_issue = time.perf_counter()
    torch.cuda.synchronize()
    t_done = time.perf_counter()
    a = acc.setdefault(name, [0.0, 0.0])
    a[0] += (t_issue - t) * 1e3
    a[1] += (t_done - t) * 1e3
    return r


acc = {}
for _ in range(N):
    h = torch.from_numpy(np.ascontiguousarray(np.stack(host_list)))
    control = stage("h2d + normalise", lambda: h.to(dev, non_blocking=True).to(torch.float32).div_(255.0).clamp_(0, 1)
                    .permute(0, 3, 1, 2).contiguous(), acc)
    cn = stage("control * 2 - 1", lambda: control * 2 - 1, acc)
    lat = stage("vae.encode", lambda: vae.encode(cn).latent_dist.mode().to(torch.float32) * vae.config.scaling_factor, acc)
    lat2 = stage("dit (generate_sample_1step)", lambda: ir.generate_sample_1step(net, sched, lat, 400, y, mask, use_control=True), acc)
    im = stage("vae.decode", lambda: vae.decode_tensor(lat2, in_scale=1.0 / 0.18215, out_scale=0.5, out_shift=0.5), acc)
    stage("uint8 + d2h", lambda: pipeline._to_host_pair(pipeline.to_uint8_nhwc(im), pipeline.to_uint8_nhwc(control)), acc)
tot = 0.0
for k, (a, b) in acc.items():
    print(f"  {k:32s} host issue {a / N:7.3f} ms   until done {b / N:7.3f} ms")
    tot += b / N
print(f"  sum of synchronised stages: {tot:.3f} ms")
