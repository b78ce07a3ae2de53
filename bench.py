#!/usr/bin/env python3
"""bench.py - LQR solves/sec of the B200-native LQR step (BASELINE.json metric).

Workload (N=1): BASELINE config 3, "Random LTI batched LQR, n_batch=4096, T=20, n_state=8,
n_ctrl=2" in fp32, unbounded; one *step* = one LQRStepFn.forward over the batch (Riccati sweep +
line-search rollout) = ONE kernel launch.  N>1: every rank owns its own 4096-problem shard
(weak scaling, no collective on the solve path; NCCL only for the barrier and the max-over-ranks).

  value      device-resident throughput: inputs in HBM, rotating over 4 input sets (281 MB > L2)
  e2e        same metric through the public API LQRStep(...)(x_init,C,c,F,f) with the inputs in
             pinned HOST memory: H2D of the step's inputs and D2H of its results inside the timing
  roofline   algorithmic bytes per launch / kernel time  vs the measured HBM copy peak
  cpu_baseline  the oracle port (oracle/lqr_oracle.py, vectorised torch CPU) on the host cores

`--impl reference` times the CPU path alone (the oracle port: the reference is pure Python and
does not travel).
"""
import argparse
import ctypes
import faulthandler
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

CFG = dict(B=4096, T=20, n=8, m=2)
WORKLOAD = "config3: random LTI batched LQR, n_batch=4096/GPU, T=20, n_state=8, n_ctrl=2, fp32, unbounded"
N_SETS = 4


def bytes_per_solve(T, n, m, tensor_bounds=False):
    """every input read once, every output written once (fp32)."""
    p = n + m
    inp = T * p * p + T * p + (T - 1) * n * p + (T - 1) * n + n + T * n + T * m
    if tensor_bounds:
        inp += 2 * T * m
    return 4 * inp + 4 * (T * n + T * m + 2)


def gen_inputs(seed, B, T, n, m, device):
    """Synthetic generator of SURVEY.md section 8(d) (same as tests/helpers.gen_problem, on device)."""
    g = torch.Generator(device=device).manual_seed(seed)
    p = n + m
    L = torch.randn(T, B, p, p, generator=g, device=device) / p ** 0.5
    C = L @ L.transpose(-1, -2) + torch.eye(p, device=device)
    c = torch.randn(T, B, p, generator=g, device=device)
    A = 0.9 * torch.eye(n, device=device) + 0.1 * torch.randn(B, n, n, generator=g, device=device) / n ** 0.5
    Bm = torch.randn(B, n, m, generator=g, device=device) / n ** 0.5
    F = torch.cat((A, Bm), -1).unsqueeze(0).repeat(T - 1, 1, 1, 1).contiguous()
    f = 0.1 * torch.randn(T - 1, B, n, generator=g, device=device)
    x0 = torch.randn(B, n, generator=g, device=device)
    u = torch.zeros(T, B, m, device=device)
    xs = [x0]
    for t in range(T - 1):
        xs.append(torch.einsum("bij,bj->bi", F[t], torch.cat((xs[t], u[t]), 1)) + f[t])
    x = torch.stack(xs)
    return dict(x_init=x0, C=C.contiguous(), c=c, F=F, f=f, cur_x=x, cur_u=u)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the GPU is under this benchmark's load."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,utilization.gpu")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([s.strip() for s in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        rows = [r for r in self.rows if len(r) >= 7 and r[0].isdigit()]
        busy = [r for r in rows if r[6].isdigit() and int(r[6]) > 0] or rows
        reasons = set()
        for r in busy:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(int(r[0]) for r in busy) if busy else None,
                "sm_max_mhz": int(busy[0][1]) if busy else None, "reasons": sorted(reasons),
                "samples": len(busy)}


class RawStepper:
    """Pre-bound C-ABI call (device-resident arm): one ctypes call = one kernel launch."""

    def __init__(self, inp, B, T, n, m):
        from mpc.pytorch_b200 import _lib
        from mpc.pytorch_b200._lib import Dims, Params, ptr
        dev = inp["C"].device
        self.out = dict(new_x=torch.empty(T, B, n, device=dev), new_u=torch.empty(T, B, m, device=dev),
                        costs=torch.empty(B, device=dev), fdn=torch.empty(B, device=dev),
                        alphas=torch.empty(B, device=dev))
        self.dims = Dims(B=B, T=T, n=n, m=m, F_T=T - 1, has_f=1, bounds_kind=0, has_zero_mask=0,
                         has_delta_u=0, max_ls_iter=10, pnqp_max_iter=20, do_rollout=1)
        self.params = Params(u_lo=0.0, u_hi=0.0, delta_u=0.0, ls_decay=0.2)
        self.fn = _lib.lib().mpcb200_lqr_step_f32
        o = self.out
        self.args = [ctypes.byref(self.dims), ctypes.byref(self.params), ptr(inp["C"]), ptr(inp["c"]),
                     ptr(inp["F"]), ptr(inp["f"]), ptr(inp["x_init"]), ptr(inp["cur_x"]), ptr(inp["cur_u"]),
                     None, None, None, ptr(o["new_x"]), ptr(o["new_u"]), ptr(o["costs"]), ptr(o["fdn"]),
                     ptr(o["alphas"]), None, None, None, None, None, None, None]
        self.keep = inp

    def __call__(self, stream):
        self.args[-1] = stream
        rc = self.fn(*self.args)
        if rc != 0:
            raise RuntimeError(f"mpcb200_lqr_step_f32 -> {rc}")


def cpu_reference_arm(steps, warmup, sample_B=None):
    """The reference's CPU path for this workload = the oracle port on all host threads."""
    from oracle import lqr_oracle as orc
    B, T, n, m = (sample_B or CFG["B"]), CFG["T"], CFG["n"], CFG["m"]
    inp = {k: v for k, v in gen_inputs(3000, B, T, n, m, torch.device("cpu")).items()}

    def one():
        return orc.lqr_step_forward(n, m, T, inp["x_init"], inp["C"], inp["c"], inp["F"], inp["f"],
                                    inp["cur_x"], inp["cur_u"], coupled=True)
    # "all the host threads it can use": the tiny batched ops stop scaling early, so calibrate the
    # thread count on one step each and keep the fastest
    ncpu = os.cpu_count() or 1
    best = None
    for cand in sorted({c for c in (4, 8, 16, 32, 64, ncpu) if c <= ncpu}):
        torch.set_num_threads(cand)
        one()
        t0 = time.perf_counter()
        one()
        dt = time.perf_counter() - t0
        if best is None or dt < best[0]:
            best = (dt, cand)
        if dt > 3 * best[0]:
            break
    cores = best[1]
    torch.set_num_threads(cores)
    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):                 # bounded sample: stop after ~20 s of CPU work
        one()
        done += 1
        if time.perf_counter() - t0 > 20.0:
            break
    dt = time.perf_counter() - t0
    return (B * done / dt, dt / done * 1e3, cores,
            f"{done} x full {WORKLOAD.split(':')[0]} batch (B={B}), oracle port on torch CPU, {cores} threads")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    a = ap.parse_args()
    # watchdog: a hung bench must not eat the box; dumps all Python stacks and exits
    faulthandler.dump_traceback_later(int(os.environ.get("BENCH_WATCHDOG_S", "900")), exit=True)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    B, T, n, m = CFG["B"], CFG["T"], CFG["n"], CFG["m"]
    bps = bytes_per_solve(T, n, m)
    base = {"metric": "LQR solves/sec", "unit": "solves/s", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "n_batch_per_gpu": B, "T": T, "n_state": n, "n_ctrl": m,
                       "parallelism": f"batch-shard x{a.gpus} (no collective on the solve path)"}}

    if a.impl == "reference":
        if rank != 0:
            return
        steps = max(1, min(a.steps, 40))
        val, ms, cores, sample = cpu_reference_arm(steps, max(1, min(a.warmup, 3)))
        base.update({"impl": "reference", "value": val, "ms_per_step": ms, "steps": steps,
                     "cpu_baseline": {"value": val, "unit": "solves/s", "cores": cores, "kind": "port",
                                      "sample": sample},
                     "e2e": {"value": val, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                     "gpu_launches": 0})
        print(json.dumps(base))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from mpc.pytorch_b200 import LQRStep, QuadCost, LinDx, _lib
    sets = [gen_inputs(1000 * 3 + rank * 17 + s, B, T, n, m, dev) for s in range(N_SETS)]
    steppers = [RawStepper(s, B, T, n, m) for s in sets]
    stream = torch.cuda.current_stream(dev)
    sh = ctypes.c_void_p(stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # pre-heat (untimed, before the W warm-up steps): ~0.7 s of the same launches so that the
    # nvidia-smi samples are taken with the SMs at their loaded clocks
    preheat = float(os.environ.get("BENCH_PREHEAT_S", "0.7"))   # set 0 under ncu (every launch is replayed)
    t_end = time.perf_counter() + preheat
    i = 0
    while time.perf_counter() < t_end:
        for _ in range(50):
            steppers[i % N_SETS](sh)
            i += 1
        torch.cuda.synchronize(dev)
    for w in range(a.warmup):
        steppers[w % N_SETS](sh)
    barrier()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(a.steps):
        steppers[k % N_SETS](sh)
    e1.record(stream)
    barrier()
    launches = _lib.launch_count() - l0
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        tms = torch.tensor([ms], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())

    # ---------------- e2e: public API, host buffers, copies inside the timed region
    # Two streams: the copy-in stream feeds double-buffered device inputs while the run stream solves the
    # previous step and copies its results out (H2D and D2H use different DMA engines).  Every step still
    # moves all of its inputs host->device and its results device->host inside the timed region.
    host = [{k: v.cpu().pin_memory() for k, v in s.items()} for s in sets[:2]]
    dbufs = [{k: torch.empty_like(v) for k, v in sets[0].items()} for _ in range(2)]
    h_outs = [[torch.empty(T, B, n).pin_memory(), torch.empty(T, B, m).pin_memory(), torch.empty(B).pin_memory()]
              for _ in range(2)]
    h2d = sum(v.numel() * v.element_size() for v in host[0].values())
    d2h = sum(v.numel() * v.element_size() for v in h_outs[0])
    s_in, s_run = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]

    def e2e_step(k):
        i = k % 2
        dbuf, h_out = dbufs[i], h_outs[i]
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_free[i])                       # the solve that last read dbufs[i] is done
            for key, v in host[i].items():
                dbuf[key].copy_(v, non_blocking=True)
            ev_in[i].record(s_in)
        with torch.cuda.stream(s_run):
            s_run.wait_event(ev_in[i])
            step = LQRStep(n, m, T, true_cost=QuadCost(dbuf["C"], dbuf["c"]),
                           true_dynamics=LinDx(dbuf["F"], dbuf["f"]),
                           current_x=dbuf["cur_x"], current_u=dbuf["cur_u"])
            nx, nu, _, costs, _, _ = step(dbuf["x_init"], dbuf["C"], dbuf["c"], dbuf["F"], dbuf["f"])
            ev_free[i].record(s_run)
            h_out[0].copy_(nx, non_blocking=True)
            h_out[1].copy_(nu, non_blocking=True)
            h_out[2].copy_(costs, non_blocking=True)

    e2e_steps = max(3, min(a.steps, 50))
    with torch.no_grad():
        for k in range(3):
            e2e_step(k)
        barrier()
        t0 = time.perf_counter()
        for k in range(e2e_steps):
            e2e_step(k)
        torch.cuda.synchronize(dev)
        e2e_s = time.perf_counter() - t0
    if world > 1:
        te = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te.item())
        dist.barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = bps * B * a.steps / (ms * 1e-3) / 1e9          # GB/s per GPU (max-over-ranks time)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("lqr_step_f32_8_2_dram_bytes_per_launch")
    out = dict(base)
    out.update({
        "value": B * world * a.steps / (ms * 1e-3), "ms_per_step": ms / a.steps,
        "gpu_launches": int(launches), "clocks": clocks,
        "e2e": {"value": B * world * e2e_steps / e2e_s, "unit": "solves/s",
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                "api": "mpc.pytorch_b200.LQRStep(...)(x_init,C,c,F,f), pinned host buffers, copy-in / run streams"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "kernel": "lqr_step_kernel<float,8,2>", "algorithmic_bytes_per_launch": bps * B,
                     "bytes_per_solve": bps},
    })
    out["config"]["l2_policy"] = f"rotating {N_SETS} input sets ({N_SETS * bps * B / 1e6:.0f} MB > 126 MB L2)"
    out["config"]["preheat_s"] = preheat
    if world == 1:
        val, _, cores, sample = cpu_reference_arm(40, 2)
        out["cpu_baseline"] = {"value": val, "unit": "solves/s", "cores": cores, "kind": "port", "sample": sample}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
