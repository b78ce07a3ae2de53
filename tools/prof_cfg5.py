"""Launch loop for ncu: the config-5 shard (B=4096, T=50, n=16, m=4) step kernel."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
dev = torch.device("cuda:0")
sets = [bench.gen_inputs(5100 + s, 4096, 50, 16, 4, dev) for s in range(2)]
st = [bench.RawStepper(s, 4096, 50, 16, 4) for s in sets]
sh = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    st[i % 2](sh)
torch.cuda.synchronize()
print("done")
