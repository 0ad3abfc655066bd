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

from tools.workloads import (WORKLOADS, ROBIN_L, ROBIN_R, build_operator, build_oracle, check_rows, field_planes)  # noqa: E402


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


def global_shape_of(args, N):
    """-> (global extents, workload description) of this run: BASELINE's workload, `--nz` override, weak scaling along dim 3."""
    shape, _, _, desc = WORKLOADS[args.workload]
    gshape = tuple(shape)
    if args.nz > 0:
        gshape = gshape[:-1] + (args.nz,)
        desc += f" [last extent overridden to {args.nz}]"
    if args.scaling == "weak" and N > 1:
        gshape = gshape[:-1] + (gshape[-1] * N,)
    return gshape, desc


def config_of(args, N, gshape, desc, es):
    """The `config` object of the JSON line: what is computed, not how -- identical for this repository's arm and for the
    reference arm of the same command line."""
    total_bytes = float(np.prod(gshape)) * es
    return {"workload": desc if args.scaling == "strong" or N == 1 else desc + f" (weak: {'x'.join(map(str, gshape))} global)",
            "global_shape": list(gshape), "parallelism": f"slab{N}" if N > 1 else "single",
            "l2": "input (>= 8 GB) far exceeds the 126 MB L2; no flush needed" if total_bytes > 1e9 else "working set may be L2-resident",
            "values": "iid Uniform(-1,1), seeded per plane"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    name = args.workload
    dtype = WORKLOADS[name][2]
    N = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    gshape, desc = global_shape_of(args, N)
    nthreads = os.cpu_count() or 1
    gpts, ms, sample = time_oracle(name, dtype, max(1, args.steps), min(args.warmup, 1), nthreads)
    line = {
        "impl": "reference", "metric": "stencil Gpoints/s (mul!, Float64)", "value": gpts, "unit": "Gpoints/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64" if dtype == np.float64 else "f32", "data": "synthetic",
        "config": config_of(args, N, gshape, desc, np.dtype(dtype).itemsize),
        "cpu_baseline": {"value": gpts, "unit": "Gpoints/s", "cores": nthreads, "kind": "port", "sample": sample,
                         "note": "reference CPU path = C restatement of DiffEqOperators.jl's mul! (oracle/), Julia is not installed on the box; "
                                 "timed on a bounded sample of the workload with all host threads, rank 0 only"},
        "e2e": {"value": gpts, "unit": "Gpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def pinned_array(L, shape_, dt):
    """Column-major numpy array over pinned host memory (deo_host_alloc); pageable fallback."""
    import ctypes as C
    from deo_b200 import _lib
    nbytes = int(np.prod(shape_)) * np.dtype(dt).itemsize
    p = C.c_void_p()
    try:
        _lib.check(L.deo_host_alloc(nbytes, C.byref(p)))
        buf = (C.c_char * nbytes).from_address(p.value)
        return np.frombuffer(buf, dtype=dt).reshape(shape_, order="F"), p
    except Exception:
        return np.empty(shape_, dtype=dt, order="F"), None


def parity_of_rows(name, gshape, dtype, du_local, start, count, nthreads, planes=8):
    """Checker leg (outside every timed region): the first / middle / last `planes` rows of this rank's slab of the TIMED
    result against the CPU oracle on the same reproducible field -> (max abs err, max |oracle|, boundary-only max abs err,
    rows checked)."""
    from oracle import oracle as O
    u_of = lambda a, b: field_planes(gshape, dtype, a, b)
    if len(gshape) == 1:
        e, r, b = check_rows(O, name, gshape, dtype, du_local, 0, gshape[0], u_of, nthreads)
        return e, r, b, int(gshape[0])
    planes = min(planes, count)
    starts = sorted({0, max((count - planes) // 2, 0), count - planes})
    err = ref = berr = 0.0
    for s0 in starts:
        e, r, b = check_rows(O, name, gshape, dtype, du_local[..., s0:s0 + planes], start + s0, start + s0 + planes, u_of, nthreads)
        err, ref, berr = max(err, e), max(ref, r), max(berr, b)
    return err, ref, berr, planes * len(starts)


def roofline_dict(name, kernel_name, alg_bytes, ms_kernel, local_frac=1.0):
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (ms_kernel * 1e-3) / 1e9
    tr = ncu_traffic(name)
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": (tr * local_frac if tr else None),
            "traffic_note": "dram bytes per launch from the committed single-GPU ncu capture (profiles/traffic.json), scaled to this rank's points",
            "peak_source": peak_src, "kernel": kernel_name,
            "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": ms_kernel,
            "frac_of_8TBs_nominal": achieved / 8000.0}


def table_bytes(G, dtype):
    """Coefficient traffic counted once per application: per-row weight tables and coefficient vectors."""
    from deo_b200.apply import _terms
    es = np.dtype(dtype).itemsize
    b = 0
    for Lop, _ in _terms(G):
        if Lop.nonuniform:
            b += (np.asarray(Lop.stencil_coefs).size + np.asarray(Lop.low_boundary_coefs).size + np.asarray(Lop.high_boundary_coefs).size) * es
        if np.ptp(np.asarray(Lop.coefficients)) != 0:
            b += Lop.len * es
    return b


def run_single_config(D, L, name, steps, warmup, device, check=True):
    """One BASELINE configuration on one GPU, kernel-only: K graph-replayed applications between CUDA events on the
    library stream, inputs resident; then the timed result is checked against the oracle (checker leg, untimed)."""
    import ctypes as C
    from deo_b200 import _lib
    shape, _, dtype, desc = WORKLOADS[name]
    es = np.dtype(dtype).itemsize
    G = build_operator(D, name, shape, dtype)
    plan = D.build_plans(G, shape, shape, dtype)[0][0]
    kernel_name, _ = plan.info
    u_host = field_planes(shape, dtype, 0, shape[-1])
    u = D.DeviceArray.from_host(u_host)
    del u_host
    du = D.DeviceArray(shape, dtype)
    reps = max(steps, 1000) if name == "C1" else steps      # C1: >= 1000 back-to-back applies from one graph
    for _ in range(max(warmup, 3)):
        plan.apply(du, u)
    D.sync()
    plan.time(du, u, reps)                                   # builds and warms the graph
    with ClockSampler(device) as clk:
        ms = plan.time(du, u, reps)
        if name == "C1":
            for _ in range(20):                              # a 3 ms region is too short for the 3 ms NVML poll: repeat it
                ms = min(ms, plan.time(du, u, reps))
    pts = float(np.prod(shape))
    alg = 2.0 * es * pts + table_bytes(G, dtype)
    out = {"config": name, "workload": desc, "shape": list(shape), "dtype": "f64" if es == 8 else "f32", "kernel": kernel_name,
           "steps": reps, "ms_per_step": ms, "value": pts / (ms * 1e-3) / 1e9, "unit": "Gpoints/s",
           "roofline": roofline_dict(name, kernel_name, alg, ms), "clocks": clk.summary(),
           "l2": "working set (16 MB) is L2-resident: the HBM fraction is nominal" if name == "C1" else "field >= 1 GB, far larger than the 126 MB L2"}
    if name == "C1":
        out["single_launch_ms"] = min(plan.time(du, u, 1) for _ in range(20))
    if check:
        got = du.to_host()
        e, r, b, rows = parity_of_rows(name, shape, dtype, got, 0, shape[-1], os.cpu_count() or 1)
        out["parity"] = {"max_rel_err": e / r, "boundary_max_rel_err": b / r, "rows_checked": rows,
                         "gate": 1e-13 if es == 8 else 1e-5, "against": "CPU oracle (oracle/), same seeded field"}
    return out


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
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config array (C1..C4) of the single-GPU line")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of the timed result")
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

    # stdout carries exactly ONE JSON line: anything a library prints there while we run (NCCL's version banner, ...) is
    # sent to stderr instead; the real stdout comes back for the final print
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

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
    shape, _, dtype, _ = WORKLOADS[name]
    if N > 1 and len(shape) != 3:
        raise SystemExit("multi-GPU runs use a 3-D workload (slabs along dim 3)")
    gshape, desc = global_shape_of(args, N)
    es = np.dtype(dtype).itemsize
    flags = _lib.DEO_FLAG_FORCE_GENERIC if args.force_generic else 0
    G = build_operator(D, name, gshape, dtype)
    total_pts = float(np.prod(gshape))

    from deo_b200.dist import SlabContext, SlabPlan
    if N > 1:
        ctx = SlabContext.from_torch_distributed(device=local_rank)
        plan = SlabPlan(G, gshape, dtype, ctx=ctx, flags=flags)
        in_shape, out_shape = plan.local_in_shape, plan.local_out_shape
        start, count, halo = plan.start, plan.count, plan.halo
    else:
        plan = D.build_plans(G, gshape, gshape, dtype, flags=flags)[0][0]
        in_shape, out_shape = gshape, gshape
        start, count, halo = 0, gshape[-1], 0
    kernel_name, launches_per_apply = plan.info
    local_pts = float(np.prod(out_shape))

    # synthetic field: iid Uniform(-1,1), reproducible plane by plane, generated on the host in pinned memory (also the
    # e2e source buffer).  A rank's buffer is [halo | own planes | halo]; its halo planes hold the neighbours' values
    # (the exchange rewrites them with the same numbers), zeros outside the physical faces.
    u_host, u_pin = pinned_array(L, in_shape, dtype)
    du_host, du_pin = pinned_array(L, out_shape, dtype)
    if len(gshape) == 1:
        u_host[...] = field_planes(gshape, dtype, 0, gshape[0])
    else:
        n_last = gshape[-1]
        u_host[...] = 0
        g0, g1 = max(start - halo, 0), min(start + count + halo, n_last)
        blk = 64
        for a in range(g0, g1, blk):                 # chunked fill keeps the temporary small
            b = min(a + blk, g1)
            u_host[..., a - (start - halo):b - (start - halo)] = field_planes(gshape, dtype, a, b)
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
        # N == 1: CUDA events on the library stream around K graph-replayed applies;
        # N > 1 : events around K applies, each = {copy-engine halo pushes || ONE fused stencil launch}
        ms_step = plan.time(du, u, args.steps)
        D.sync()
        barrier()
    ms_step = max_over_ranks(ms_step)
    launches = D.launch_count() - launches0
    clocks = clk.summary()
    value = total_pts / (ms_step * 1e-3) / 1e9

    # ---- roofline of the dominant kernel: one launch per step (N > 1: the fused slab launch, which includes its
    # in-kernel wait for the halo planes, so ms_per_launch == ms_per_step there) ---------------------------------
    alg_bytes = 2.0 * es * local_pts + table_bytes(G, dtype)
    roofline = roofline_dict(name, kernel_name, alg_bytes, ms_step, local_pts / float(np.prod(WORKLOADS[name][0])))
    if N > 1:
        roofline["note"] = "per rank: this rank's algorithmic bytes over the step time (halo exchange and in-kernel halo wait included)"

    # ---- parity of the TIMED result against the oracle (untimed checker leg) -----------------------------------
    parity = None
    if not args.no_parity and not args.force_generic:
        _lib.check(L.deo_buffer_download(du_host.ctypes.data_as(C.c_void_p), du._h, du.nbytes))
        nthreads = max(1, (os.cpu_count() or 1) // N)
        e, r, b, rows = parity_of_rows(name, gshape, dtype, du_host, start, count, nthreads)
        rel, brel = max_over_ranks(e / r), max_over_ranks(b / r)
        parity = {"max_rel_err": rel, "boundary_max_rel_err": brel, "rows_checked_per_rank": rows, "ranks": N,
                  "gate": 1e-13 if es == 8 else 1e-5, "pass": bool(rel <= (1e-13 if es == 8 else 1e-5)),
                  "against": "CPU oracle (oracle/) on the same reproducible field: first / middle / last planes of every rank's slab"}

    # ---- end to end through the public host-buffer API: H2D(u) + apply + D2H(du) every step ------------------
    e2e_steps = max(1, args.e2e_steps)
    h2d = int(np.prod(in_shape)) * es
    d2h = int(np.prod(out_shape)) * es

    def e2e_step():
        if N == 1:
            D.mul_(du_host, G, u_host, flags=flags)                 # deo_plan_apply_host: upload, fused kernel, download
        else:
            plan.apply_host(du_host, u_host)                        # deo_dist_plan_apply_host: the same pipeline on a slab
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

    # ---- the other BASELINE configurations, kernel-only with their own oracle check (N == 1 only) -------------------
    configs = None
    if N == 1 and not args.no_configs and name == "C5" and args.nz == 0 and not args.force_generic:
        del u, du
        configs = []
        for cname in ("C1", "C2", "C3", "C3f32", "C4"):
            try:
                configs.append(run_single_config(D, L, cname, args.steps, args.warmup, local_rank, check=not args.no_parity))
            except Exception as exc:          # a failing side config must not hide the headline line
                configs.append({"config": cname, "error": repr(exc)})

    if rank == 0:
        line = {
            "metric": "stencil Gpoints/s (mul!, Float64)" if dtype == np.float64 else "stencil Gpoints/s (mul!, Float32)",
            "value": value, "unit": "Gpoints/s", "n_gpus": N, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64" if dtype == np.float64 else "f32", "data": "synthetic",
            "config": config_of(args, N, gshape, desc, es), "kernel": kernel_name,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
        }
        if parity is not None:
            line["parity"] = parity
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if configs is not None:
            line["configs"] = configs
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    for p in (u_pin, du_pin):
        if p is not None:
            L.deo_host_free(p)
    return 0


if __name__ == "__main__":
    sys.exit(main())
