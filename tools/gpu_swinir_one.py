import sys, torch
sys.path.insert(0, "/root/repo")
import instarevive_b200 as ir
from instarevive_b200 import weights
dev = torch.device("cuda:0")
net = ir.SwinIR(weights.make_swinir_state_dict(seed=7), device=dev)
x = torch.rand(1, 3, 1024, 1024, device=dev)
net(x); torch.cuda.synchronize()
torch.cuda.profiler.start(); net(x); torch.cuda.synchronize(); torch.cuda.profiler.stop()
