#!/usr/bin/env python3
"""Minimal launch loop for ncu: N launches of the config-3 step kernel over rotating inputs.
usage: prof_driver.py [n_launches] [bounded:0|1] [B]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

n_launch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
B = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
dev = torch.device("cuda:0")
sets = [bench.gen_inputs(3000 + s, B, 20, 8, 2, dev) for s in range(4)]
st = [bench.RawStepper(s, B, 20, 8, 2) for s in sets]
if len(sys.argv) > 2 and sys.argv[2] == "1":
    for s in st:
        s.dims.bounds_kind = 1
        s.params.u_lo, s.params.u_hi = -0.25, 0.25
sh = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
for i in range(n_launch):
    st[i % 4](sh)
torch.cuda.synchronize()
print("done", n_launch)
