"""Launch loop for ncu: the one-call KKT adjoint at config-3 size (prep + fused kernel)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
dev = torch.device("cuda:0")
sets = [bench.gen_inputs(3100 + s, 4096, 20, 8, 2, dev) for s in range(4)]
raws = [bench.RawAdjoint(s, s["cur_x"], s["cur_u"], 4096, 20, 8, 2) for s in sets]
sh = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 8):
    raws[i % 4](sh)
torch.cuda.synchronize()
print("done")
