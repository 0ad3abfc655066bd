#!/usr/bin/env python
"""bench.py -- headline benchmark of the fused operator application (`mul!`), BASELINE.json's metric:
stencil Gpoints/s (Float64) and fraction of the HBM roofline on 1/2/4/8 B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload C5|C2|C3|C3f32|C1] [--scaling strong|weak]

Workload (default, every N): BASELINE config 5 -- 3-D Laplacian Dxx+Dyy+Dzz, CenteredDifference(2,4),
RobinBC on every face, 1024^3 Float64, slab-decomposed along dim 3 over the N ranks (strong scaling;
`--scaling weak` keeps 1024^3 per rank).  It fits one GPU (8.6 GB field + 8.6 GB result), so N=1 runs the
same job.  A "step" is one fused application over the whole grid.

One JSON line on stdout (rank 0).  `value` = Gpoints/s with u resident in HBM; `e2e` = the same through
the host-buffer form of mul! (H2D of u and D2H of du inside the timed region); `roofline` = algorithmic
bytes (16 B/point F64: one read of u, one write of du) over the kernel's CUDA-event duration, against the
measured HBM copy bandwidth in MEASURED_PEAKS.json; `cpu_baseline` = the CPU oracle (a C restatement of
the reference's CPU path; Julia is not available) on a bounded sample, on the box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (shape, approximation order, dtype, description)
    "C5": ((1024, 1024, 1024), 4, np.float64, "C5: 3-D Laplacian Dxx+Dyy+Dzz CenteredDifference(2,4) + RobinBC, 1024^3 Float64, slabs along dim 3"),
    "C3": ((512, 512, 512), 6, np.float64, "C3: 3-D Laplacian (2,6) + Neumann MultiDimBC, 512^3 Float64"),
    "C3f32": ((512, 512, 512), 6, np.float32, "C3: 3-D Laplacian (2,6) + Neumann MultiDimBC, 512^3 Float32"),
    "C2": ((8192, 8192), 4, np.float64, "C2: 2-D Laplacian Dxx+Dyy (2,4) + RobinBC, 8192^2 Float64"),
    "C1": ((10 ** 6,), 2, np.float64, "C1: 1-D heat-equation Laplacian CenteredDifference(2,2)*Dirichlet0BC, N=1e6 Float64"),
    "C4": ((512, 512, 512), 4, np.float64, "C4: non-uniform 512^3 Float64, sum over axes of CenteredDifference(2,4) + CenteredDifference(1,4) + "
           "UpwindDifference(1,2) with a mixed-sign coefficient vector, RobinBC from the same spacings (9 operators, one pass)"),
}
ROBIN_L, ROBIN_R = (1.0, 0.5, 0.25), (1.0, -0.5, 0.75)


def c4_inputs(shape, dtype):
    """Spacing vectors and coefficient vectors of BASELINE config 4."""
    hs = [1.0 / (s + 1) for s in shape]
    dxs = [(h * (1 + 0.3 * np.sin(2 * np.pi * np.arange(1, s + 2) / (s + 1)))).astype(dtype) for s, h in zip(shape, hs)]
    cs = [np.sin(6 * np.pi * np.arange(1, s + 1) / s).astype(dtype) for s in shape]
    return dxs, cs


def build_operator(D, name, shape, dtype):
    """The product-side operator A*Q of a workload (host mirror of the reference constructors)."""
    _, a, _, _ = WORKLOADS[name]
    nd = len(shape)
    h = tuple(1.0 / (s + 1) for s in shape)
    if name == "C4":
        dxs, cs = c4_inputs(shape, dtype)
        ops = [D.CenteredDifference[ax](2, 4, dxs[ax - 1], shape[ax - 1], dtype=dtype) for ax in range(1, nd + 1)] + \
              [D.CenteredDifference[ax](1, 4, dxs[ax - 1], shape[ax - 1], dtype=dtype) for ax in range(1, nd + 1)] + \
              [D.UpwindDifference[ax](1, 2, dxs[ax - 1], shape[ax - 1], cs[ax - 1], dtype=dtype) for ax in range(1, nd + 1)]
        A = ops[0]
        for o in ops[1:]:
            A = A + o
        return A * D.compose(*D.RobinBC(ROBIN_L, ROBIN_R, dxs, 1, shape, dtype=dtype))
    if nd == 1:
        return D.CenteredDifference(2, a, h[0], shape[0], dtype=dtype) * D.Dirichlet0BC(dtype)
    A = D.CenteredDifference[1](2, a, h[0], shape[0], dtype=dtype)
    for ax in range(2, nd + 1):
        A = A + D.CenteredDifference[ax](2, a, h[ax - 1], shape[ax - 1], dtype=dtype)
    if name.startswith("C3"):
        Q = D.compose(*D.Neumann0BC(dtype, h, 1, shape))
    else:
        Q = D.compose(*D.RobinBC(ROBIN_L, ROBIN_R, h, 1, shape, dtype=dtype))
    return A * Q


def build_oracle(O, name, shape, dtype):
    _, a, _, _ = WORKLOADS[name]
    nd = len(shape)
    h = tuple(1.0 / (s + 1) for s in shape)
    if name == "C4":
        dxs, cs = c4_inputs(shape, dtype)
        ops = [O.CenteredDifference(2, 4, dxs[ax], shape[ax], axis=ax + 1, dtype=dtype) for ax in range(nd)] + \
              [O.CenteredDifference(1, 4, dxs[ax], shape[ax], axis=ax + 1, dtype=dtype) for ax in range(nd)] + \
              [O.UpwindDifference(1, 2, dxs[ax], shape[ax], cs[ax], axis=ax + 1, dtype=dtype) for ax in range(nd)]
        return ops, {ax + 1: O.RobinBC(ROBIN_L, ROBIN_R, dxs[ax], 1, dtype) for ax in range(nd)}
    ops = [O.CenteredDifference(2, a, h[ax], shape[ax], axis=ax + 1, dtype=dtype) for ax in range(nd)]
    if nd == 1:
        bcs = {1: O.Dirichlet0BC(dtype)}
    elif name.startswith("C3"):
        bcs = {ax + 1: O.Neumann0BC(h[ax], 1, dtype) for ax in range(nd)}
    else:
        bcs = {ax + 1: O.RobinBC(ROBIN_L, ROBIN_R, h[ax], 1, dtype) for ax in range(nd)}
    return ops, bcs


def cpu_sample_shape(shape):
    """Bounded sample of the workload for the CPU legs: same operator, same x-y extent where possible,
    truncated along the last axis to ~3.4e7 points (about 10-30 s of single-core work)."""
    if len(shape) == 1:
        return shape
    budget = 2 ** 25
    s = list(shape)
    while np.prod(s) > budget and s[-1] > 32:
        s[-1] //= 2
    while np.prod(s) > budget and s[0] > 64:
        s[0] //= 2
    return tuple(int(v) for v in s)


def time_oracle(name, dtype, steps, warmup, nthreads):
    """Gpoints/s of the CPU oracle (C restatement of the reference's mul! path) on the bounded sample."""
    from oracle import oracle as O
    full = WORKLOADS[name][0]
    shape = cpu_sample_shape(full)
    ops, bcs = build_oracle(O, name, shape, dtype)
    u = np.asfortranarray(np.random.default_rng(0).uniform(-1, 1, shape).astype(dtype))
    for _ in range(max(0, warmup)):
        O.apply_sum(ops, u, bcs, nthreads=nthreads)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.apply_sum(ops, u, bcs, nthreads=nthreads)
    dt = (time.perf_counter() - t0) / steps
    pts = float(np.prod(shape))
    sample = f"{'x'.join(map(str, shape))} block of the {'x'.join(map(str, full))} workload, same operator and BCs, {steps} applications"
    return pts / dt / 1e9, dt * 1e3, sample


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line).  NVML is polled from a
    thread every few milliseconds (the timed region is tens of milliseconds, too short for `nvidia-smi -lms`); if NVML
    is not importable the nvidia-smi loop is used instead."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.samples = []          # (sm_mhz, max_mhz, reasons bitmask or set)
        self.proc = None
        self.stop = threading.Event()
        self.thread = None
        self.nvml = None

    def _index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.device])
            except Exception:
                return self.device
        return self.device

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._index())
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return self
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self._index())], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read_smi, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _poll_nvml(self):
        n = self.nvml
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(n, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self.stop.is_set():
            try:
                self.samples.append((float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)), self.max_mhz, int(get_reasons(self.handle))))
            except Exception:
                pass
            time.sleep(0.003)

    def _read_smi(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.proc.stdout:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                self.samples.append((float(f[1]), float(f[2]), {nm for nm, val in zip(names, f[5:9]) if val.lower().startswith("active")}))
            except ValueError:
                continue

    def __exit__(self, *exc):
        self.stop.set()
        if self.proc is not None:
            time.sleep(0.05)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        if self.thread is not None:
            self.thread.join(timeout=2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = [s[0] for s in self.samples]
        reasons = set()
        for s in self.samples:
            if isinstance(s[2], set):
                reasons |= s[2]
            else:
                bits = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
                reasons |= {nm for b, nm in bits.items() if s[2] & b}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(s[1] for s in self.samples)), "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md, MEASURED_PEAKS.json absent)"


def ncu_traffic(name):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if there is one."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(name)
    except Exception:
        return None


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    name = args.workload
    dtype = WORKLOADS[name][2]
    nthreads = os.cpu_count() or 1
    gpts, ms, sample = time_oracle(name, dtype, max(1, args.steps), min(args.warmup, 1), nthreads)
    line = {
        "impl": "reference", "metric": "stencil Gpoints/s (mul!, Float64)", "value": gpts, "unit": "Gpoints/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64" if dtype == np.float64 else "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[name][3], "note": "reference CPU path = C restatement of DiffEqOperators.jl's mul! (oracle/), "
                   "Julia is not installed on the box; timed on a bounded sample with all host threads"},
        "cpu_baseline": {"value": gpts, "unit": "Gpoints/s", "cores": nthreads, "kind": "port", "sample": sample},
        "e2e": {"value": gpts, "unit": "Gpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C5", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--nz", type=int, default=0, help="override the last extent of the workload (experiments; not a BASELINE config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--force-generic", action="store_true", help="run the per-point kernel instead of the tiled one (A/B)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        print(f"[bench] WORLD_SIZE={world} != --gpus {args.gpus}; using WORLD_SIZE", file=sys.stderr)
    N = world

    import ctypes as C
    import deo_b200 as D
    from deo_b200 import _lib
    L = D.load_library()
    _lib.check(L.deo_init(local_rank))

    dist = None
    if N > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")    # keep stdout to the one JSON line if NCCL_DEBUG is set
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    name = args.workload
    shape, _, dtype, desc = WORKLOADS[name]
    if N > 1 and len(shape) != 3:
        raise SystemExit("multi-GPU runs use a 3-D workload (slabs along dim 3)")
    gshape = tuple(shape)
    if args.nz > 0:
        gshape = gshape[:-1] + (args.nz,)
        desc += f" [last extent overridden to {args.nz}]"
    if args.scaling == "weak" and N > 1:
        gshape = gshape[:-1] + (gshape[-1] * N,)
    es = np.dtype(dtype).itemsize
    flags = _lib.DEO_FLAG_FORCE_GENERIC if args.force_generic else 0
    G = build_operator(D, name, gshape, dtype)
    total_pts = float(np.prod(gshape))

    from deo_b200.dist import SlabContext, SlabPlan
    rng = np.random.default_rng(1234 + rank)
    if N > 1:
        ctx = SlabContext.from_torch_distributed(device=local_rank)
        plan = SlabPlan(G, gshape, dtype, ctx=ctx, flags=flags)
        in_shape, out_shape = plan.local_in_shape, plan.local_out_shape
    else:
        plan = D.build_plans(G, gshape, gshape, dtype, flags=flags)[0][0]
        in_shape, out_shape = gshape, gshape
    kernel_name, launches_per_apply = plan.info
    local_pts = float(np.prod(out_shape))

    # synthetic field: iid Uniform(-1,1), generated on the host in pinned memory (also the e2e source buffer)
    def pinned(shape_, dt):
        nbytes = int(np.prod(shape_)) * np.dtype(dt).itemsize
        p = C.c_void_p()
        try:
            _lib.check(L.deo_host_alloc(nbytes, C.byref(p)))
            buf = (C.c_char * nbytes).from_address(p.value)
            return np.frombuffer(buf, dtype=dt).reshape(shape_, order="F"), p
        except Exception:
            return np.empty(shape_, dtype=dt, order="F"), None

    u_host, u_pin = pinned(in_shape, dtype)
    du_host, du_pin = pinned(out_shape, dtype)
    flat = u_host.reshape(-1, order="F")
    chunk = 1 << 24
    for i in range(0, flat.size, chunk):          # chunked fill keeps the temporary small
        flat[i:i + chunk] = rng.uniform(-1, 1, min(chunk, flat.size - i)).astype(dtype)
    u = D.DeviceArray(in_shape, dtype)
    du = D.DeviceArray(out_shape, dtype)
    _lib.check(L.deo_buffer_upload(u._h, u_host.ctypes.data_as(C.c_void_p), u.nbytes))

    def step():
        plan.apply(du, u)

    # ---- warm-up, then K timed steps between barriers; device time = max over ranks -------------------
    for _ in range(args.warmup):
        step()
    D.sync()
    barrier()
    launches0 = D.launch_count()
    with ClockSampler(local_rank) as clk:
        barrier()
        D.sync()
        if N == 1:
            ms_step = plan.time(du, u, args.steps)            # CUDA events on the library stream around K graph-replayed applies
        else:
            ms_step = plan.time(du, u, args.steps)            # events around K {exchange || interior, boundary} applies
        D.sync()
        barrier()
    ms_step = max_over_ranks(ms_step)
    launches = D.launch_count() - launches0
    clocks = clk.summary()
    value = total_pts / (ms_step * 1e-3) / 1e9

    # ---- kernel-only roofline (the dominant kernel timed alone, same events) --------------------------------
    peak, peak_src = measured_peak()
    ms_kernel = ms_step if N == 1 else max_over_ranks(plan.time(du, u, max(3, args.steps // 2)))
    alg_bytes = 2.0 * es * local_pts
    achieved = alg_bytes / (ms_kernel * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": (ncu_traffic(name) * local_pts / float(np.prod(WORKLOADS[name][0])) if ncu_traffic(name) else None),
                "traffic_note": "dram bytes per launch from the committed single-GPU ncu capture (profiles/traffic.json), scaled to this rank's points",
                "peak_source": peak_src, "kernel": kernel_name,
                "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": ms_kernel,
                "frac_of_8TBs_nominal": achieved / 8000.0}

    # ---- end to end through the public host-buffer API: H2D(u) + apply + D2H(du) every step ------------------
    e2e_steps = max(1, args.e2e_steps)
    h2d = int(np.prod(in_shape)) * es
    d2h = int(np.prod(out_shape)) * es

    def e2e_step():
        if N == 1:
            D.mul_(du_host, G, u_host, flags=flags)                 # deo_plan_apply_host: upload, fused kernel, download
        else:
            _lib.check(L.deo_buffer_upload(u._h, u_host.ctypes.data_as(C.c_void_p), h2d))
            plan.apply(du, u)
            _lib.check(L.deo_buffer_download(du_host.ctypes.data_as(C.c_void_p), du._h, d2h))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    D.sync()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / e2e_steps)
    e2e = {"value": total_pts / (e2e_ms * 1e-3) / 1e9, "unit": "Gpoints/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "ms_per_step": e2e_ms, "steps": e2e_steps, "host_buffers": "pinned" if u_pin is not None else "pageable"}

    # ---- CPU baseline: the oracle on the host cores, bounded sample, rank 0 at N == 1 only ---------------------
    cpu = None
    if rank == 0 and N == 1 and not args.no_cpu_baseline:
        g1, _, sample = time_oracle(name, dtype, 1, 0, 1)
        cpu = {"value": g1, "unit": "Gpoints/s", "cores": 1, "kind": "port", "sample": sample,
               "note": "C restatement of the reference's CPU mul! (Julia unavailable); the reference itself is single-threaded"}
        nthreads = os.cpu_count() or 1
        if nthreads > 1:
            gN, _, _ = time_oracle(name, dtype, 2, 0, nthreads)
            cpu["all_threads"] = {"value": gN, "cores": nthreads}

    if rank == 0:
        line = {
            "metric": "stencil Gpoints/s (mul!, Float64)" if dtype == np.float64 else "stencil Gpoints/s (mul!, Float32)",
            "value": value, "unit": "Gpoints/s", "n_gpus": N, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64" if dtype == np.float64 else "f32", "data": "synthetic",
            "config": {"workload": desc if args.scaling == "strong" or N == 1 else desc + f" (weak: {'x'.join(map(str, gshape))} global)",
                       "global_shape": list(gshape), "parallelism": f"slab{N}" if N > 1 else "single", "kernel": kernel_name,
                       "l2": "input (>= 8 GB) far exceeds the 126 MB L2; no flush needed" if total_pts * es > 1e9 else "working set may be L2-resident",
                       "values": "iid Uniform(-1,1), seeded"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    for p in (u_pin, du_pin):
        if p is not None:
            L.deo_host_free(p)
    return 0


if __name__ == "__main__":
    sys.exit(main())
