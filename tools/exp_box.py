import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from mpc.pytorch_b200.step import lqr_step_raw
dev = torch.device("cuda:0")
B=4096
for s in range(4):
    inp = bench.gen_inputs(3000 + s, B, 20, 8, 2, dev)
    o = lqr_step_raw(8, 2, 20, inp["x_init"], inp["C"], inp["c"], inp["F"], inp["f"], inp["cur_x"], inp["cur_u"], u_lower=-0.25, u_upper=0.25)
    torch.cuda.synchronize()
    st = bench.RawStepper(inp, B, 20, 8, 2); st.dims.bounds_kind = 1; st.params.u_lo, st.params.u_hi = -0.25, 0.25
    sh = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    for i in range(5): st(sh)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for i in range(20): st(sh)
    e1.record(); torch.cuda.synchronize()
    print(f"set {s}: {e0.elapsed_time(e1)/20*1e3:.1f} us  status!=0: {int((o['status']!=0).sum())}  qp_iters max {int(o['qp_iters'].max())} mean {float(o['qp_iters'].float().mean()):.2f}  alphas<1: {int((o['alphas']<1).sum())} min alpha {float(o['alphas'].min()):.3g}")
# isolated launches (sync between) vs back-to-back, set 0
inp = bench.gen_inputs(3000, B, 20, 8, 2, dev)
st = bench.RawStepper(inp, B, 20, 8, 2); st.dims.bounds_kind = 1; st.params.u_lo, st.params.u_hi = -0.25, 0.25
ts=[]
for i in range(10):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); st(sh); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1)*1e3)
print("isolated launches us:", [f"{t:.0f}" for t in ts])
st2 = bench.RawStepper(inp, B, 20, 8, 2)
ts=[]
for i in range(6):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); st2(sh); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1)*1e3)
print("isolated unbounded us:", [f"{t:.0f}" for t in ts])
