"""A few C2 steps (bench.section_c2's workload) for an ncu launch list:
   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c2_launches.csv python tools/c2_profile.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from jmac_b200 import _lib

bench.wall_time = lambda fn, warmup, iters: ([fn() for _ in range(2)], torch.cuda.synchronize(), 1.0)[2]   # two calls per measurement
torch.cuda.set_device(0)
torch.cuda.nvtx.range_push("c2")
out = bench.section_c2(torch.device("cuda", 0), 0, _lib.load(), graphed=False)
torch.cuda.nvtx.range_pop()
print({k: out[k] for k in ("jmac_kernel_launches_per_step", "first_step_loss")})
