import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from mpc.pytorch_b200.step import lqr_step_raw
dev = torch.device("cuda:0")
for (B, T, n, m) in [(4096, 50, 16, 4), (16384, 50, 16, 4), (4096, 20, 12, 4)]:
    s = bench.gen_inputs(7, B, T, n, m, dev)
    run = lambda: lqr_step_raw(n, m, T, s["x_init"], s["C"], s["c"], s["F"], s["f"], s["cur_x"], s["cur_u"], want_stats=False)
    for _ in range(3): run()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 10 * 1e3
    print(f"lib={os.environ.get('MPCB200_LIB','default')[-12:]} B={B} T={T} n={n} m={m}: {us:.0f} us  {B/us:.2f} Msolves/s")
