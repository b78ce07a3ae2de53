"""Per-phase clock64 report of the column-pair step kernel (needs the -DMPCB2_TIMING build):
   make -C mpc/pytorch_b200/csrc BUILD=build_timing OUT=../libmpcb200_timing.so EXTRA=-DMPCB2_TIMING
   MPCB200_LIB=mpc/pytorch_b200/libmpcb200_timing.so python tools/exp_timing.py [box]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
dev = torch.device("cuda:0")
B, T, n, m = 4096, 20, 8, 2
os.environ["MPCB200_KERNEL"] = "2"
inp = bench.gen_inputs(3000, B, T, n, m, dev)
st = bench.RawStepper(inp, B, T, n, m)
if len(sys.argv) > 1 and sys.argv[1] == "box":
    st.dims.bounds_kind = 1
    st.params.u_lo, st.params.u_hi = -0.25, 0.25
sh = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
st(sh)
torch.cuda.synchronize()
