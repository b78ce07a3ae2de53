"""n=16 step kernel: generic vs column-pair (the library chosen by MPCB200_LIB: with / without the mma.sync products)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
dev = torch.device("cuda:0")
sh = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
for (n, m, T, B, bounds) in [(16, 4, 50, 4096, None), (16, 4, 50, 16384, None), (16, 4, 50, 4096, 0.25), (16, 4, 50, 32768, None)]:
    bps = bench.bytes_per_solve(T, n, m)
    nsets = max(2, min(4, int(300e6 // (bps * B)) + 1))
    sets = [bench.gen_inputs(100 + s, B, T, n, m, dev) for s in range(nsets)]
    res, outs = {}, {}
    for impl in ("1", "2"):
        if impl == "1" and B > 4096:
            continue
        os.environ["MPCB200_KERNEL"] = impl
        sts = [bench.RawStepper(s, B, T, n, m, bounds=bounds) for s in sets]
        res[impl] = bench.time_launches(sts, 8, torch.cuda.current_stream(dev), sh)
        sts[0](sh); torch.cuda.synchronize()
        outs[impl] = {k: v.clone() for k, v in sts[0].out.items()}
    d = max(float((outs["1"][k] - outs["2"][k]).abs().max()) for k in ("new_x", "new_u")) if len(outs) == 2 else float("nan")
    print(f"lib={os.path.basename(os.environ.get('MPCB200_LIB', 'default'))} n={n} m={m} T={T} B={B} bounds={bounds}: "
          f"generic {res.get('1', float('nan')):.1f} us  pair {res['2']:.1f} us  max|generic-pair|={d:.1e}", flush=True)
    del sets, sts
    torch.cuda.empty_cache()
