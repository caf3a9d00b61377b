#!/usr/bin/env python
"""bench.py -- headline benchmark of the multigrid V-cycle hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA kernels through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores (oracle port)

Workload (BASELINE.json configs[2], the config the metric's roofline target is quoted on):
2-D Poisson, 16385 x 16385 unknowns, manufactured f = 2 pi^2 sin(pi x) sin(pi y), u0 = 0,
mixed-precision V(2,2) red-black Gauss-Seidel: fp64 iterate and residual, fp32 V-cycle correction,
switch to fp64 cycles at ||r|| <= 1e-6 (precision_strategy='adaptive').

A STEP is one multigrid cycle of the real solve loop, including its convergence test (one 8-byte
device->host read).  When a solve converges inside the timed region the next one starts from u = 0.
value = nx*ny*K / time  [fine-grid unknowns/s per V-cycle]  (reference gpu/gpu_benchmark.py:248).

Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement" for every key."""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

def _metric_name():
    try:
        return json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"]
    except Exception:
        return "V-cycle fine-grid unknowns/sec + smoother HBM GB/s vs peak at 1/2/4/8 B200"


METRIC = _metric_name()
UNIT = "unknowns/s"
BYTES_PER_UNKNOWN_FP64_V22 = 90.7  # fused-minimum traffic of one fp64 V(2,2) cycle


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", "--size", dest="n", type=int, default=16385,
                    help="fine grid points per side (per GPU slab height for N>1); use --size under torchrun, whose "
                         "own parser rejects --n as an ambiguous abbreviation")
    ap.add_argument("--strategy", default="adaptive")
    ap.add_argument("--cycle", default="V")
    ap.add_argument("--loader", default="tma")
    ap.add_argument("--smoother", default="rbgs", choices=["rbgs", "jacobi"],
                    help="1-GPU arm: red-black GS (the BASELINE config) or damped Jacobi (omega = 2/3), both in the streaming kernel")
    ap.add_argument("--tolerance", type=float, default=None,
                    help="absolute h-scaled L2 residual tolerance; default 1e-8 (the reference's).  Where that lies below "
                         "the fp64 rounding floor of f - A u (~3e-8 at h = 1/16384) the solve ends on the floor rule of "
                         "solvers/policy.py: one cycle after the residual stops contracting")
    ap.add_argument("--cpu-n", type=int, default=None,
                    help="grid of the CPU sample; default: 8193 for the cpu_baseline leg of the GPU arm (bounded to "
                         "~10 s), the configured --n for --impl reference when the host has the memory for it")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--agg", type=int, default=None,
                    help="multi-GPU: agglomerate levels with <= this many points per side (default: distributed.BENCH_AGG)")
    ap.add_argument("--halo", default="nccl", choices=["nccl", "p2p"],
                    help="multi-GPU ghost exchange: grouped NCCL send/recv (default) or one-sided pushes over NVLink peer "
                         "memory (halo.py; opt-in until measured)")
    ap.add_argument("--ghost", type=int, default=None,
                    help="multi-GPU: ghost rows per interior side (even, >= 6; default: distributed.BENCH_GHOST).  Deeper "
                         "ghosts = fewer halo exchanges per cycle (every pass spends 2 rows of validity per sweep)")
    ap.add_argument("--strong", action="store_true",
                    help="multi-GPU: --n is the GLOBAL grid (n x n on the unit square) split into row slabs")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--varcoef", action="store_true",
                    help="1-GPU arm: the variable-coefficient operator -div(a grad u), a = 1 + 0.5 sin(2 pi x) cos(pi y) + x y "
                         "(the operator of BASELINE configs[4]) through the mg_vcv_* passes; use with --n 8193")
    ap.add_argument("--no-graphs", action="store_true", help="launch every kernel eagerly instead of replaying CUDA graphs")
    ap.add_argument("--no-dd", action="store_true", help="A/B: separate defect pass and down pass (two launches) "
                                                         "instead of the fused defect + down pass")
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock, power and throttle reasons sampled DURING the timed region.  The timed region of the default run
    is tens of milliseconds, shorter than one period of `nvidia-smi -lms 200`, so the samples come from NVML directly
    (nvidia_ml_py) on a 5 ms thread and carry timestamps; `mark()` brackets the timed window and the summary is
    taken over the samples inside it.  Falls back to the nvidia-smi loop when NVML cannot be loaded."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int, period_s: float = 0.005):
        self.gpu, self.period = gpu_index, period_s
        self.proc, self.path, self.thread = None, None, None
        self.samples = []          # (t, sm_mhz, power_w, reasons_bitmask)
        self.window = [None, None]
        self._stop = False
        self.sm_max = None

    def _nvml_loop(self, h, nv):
        while not self._stop:
            try:
                self.samples.append((time.perf_counter(), nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM),
                                     nv.nvmlDeviceGetPowerUsage(h) / 1000.0,
                                     nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                                     if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons")
                                     else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
            except Exception:
                pass
            time.sleep(self.period)

    def __enter__(self):
        try:
            import threading

            import pynvml as nv
            nv.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = self.gpu
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                idx = int(vis.split(",")[self.gpu])
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self._nv = nv
            self.thread = threading.Thread(target=self._nvml_loop, args=(h, nv), daemon=True)
            self.thread.start()
            return self
        except Exception:
            self.thread = None
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def mark(self, which: int) -> None:
        """mark(0) right before the timed region starts, mark(1) right after it ends."""
        self.window[which] = time.perf_counter()

    def __exit__(self, *a):
        if self.thread is not None:
            self._stop = True
            self.thread.join(timeout=2)
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.thread is not None:
            nv = self._nv
            t0, t1 = self.window
            inside = [s for s in self.samples if t0 is not None and t1 is not None and t0 <= s[0] <= t1]
            use = inside if len(inside) >= 3 else self.samples
            out["source"] = "nvml, %g ms period; %d of %d samples inside the timed window%s" % (
                self.period * 1e3, len(inside), len(self.samples), "" if use is inside else " (too few: all samples used)")
            if use:
                sm = sorted(s[1] for s in use)
                out.update({"samples": len(use), "sm_mhz": float(sm[len(sm) // 2]), "sm_max_mhz": self.sm_max,
                            "power_w_max": max(s[2] for s in use)})
                bits = 0
                for s in use:
                    bits |= int(s[3])
                names = {"hw_slowdown": "HwSlowdown", "hw_thermal_slowdown": "HwThermalSlowdown",
                         "sw_thermal_slowdown": "SwThermalSlowdown", "sw_power_cap": "SwPowerCap"}
                for key, suffix in names.items():
                    mask = getattr(nv, "nvmlClocksEventReason" + suffix, None) or getattr(nv, "nvmlClocksThrottleReason" + suffix, 0)
                    if bits & mask:
                        out["reasons"].append(key)
            return out
        try:
            rows = [l.strip().split(", ") for l in open(self.path) if l.strip()]
            sm = sorted(float(r[1]) for r in rows)
            out["samples"] = len(rows)
            out["source"] = "nvidia-smi -lms 200"
            if sm:
                out["sm_mhz"] = sm[len(sm) // 2]
                out["sm_max_mhz"] = float(rows[0][2])
                out["power_w_max"] = max(float(r[3]) for r in rows)
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                out["reasons"] = [n for k, n in enumerate(names) if any(r[4 + k].strip() == "Active" for r in rows)]
            os.unlink(self.path)
        except Exception:
            pass
        return out


# ------------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm on the host cores (oracle port; the reference itself is pure-Python
# loops at ~6e4 unknowns/s and does not exist on the GPU box)
# ------------------------------------------------------------------------------------------------------
def cpu_cycles(n: int, cycles: int, threads_hint=None, warmup: int = 0):
    """Time `cycles` V(2,2) RB-GS fp64 cycles of the C/OpenMP oracle on an n x n grid, after `warmup` untimed cycles
    (first-touch page faults, thread-pool start).  Returns (unknowns_per_s, seconds, threads_used, residual_history)."""
    import numpy as np

    from oracle import c_oracle as CO
    from oracle import np_oracle as O
    L = 1
    a = n
    while (a - 1) % 2 == 0 and (a - 1) // 2 + 1 >= 5:
        a, L = (a - 1) // 2 + 1, L + 1
    f = O.mms_rhs(n)
    if warmup > 0:
        O.OracleMultigrid(n, max_levels=L, max_iterations=warmup, tolerance=0.0, ops=CO).solve(f)
    s = O.OracleMultigrid(n, max_levels=L, max_iterations=cycles, tolerance=0.0, ops=CO)
    t0 = time.perf_counter()
    _, info = s.solve(f)
    dt = time.perf_counter() - t0
    return n * n * info["iterations"] / dt, dt, CO.num_threads(), info["residual_history"]


def pick_cpu_threads():
    """OpenMP helps only when the cores are really there (containers often expose more CPUs than their
    quota): time one small cycle with all threads and with one, keep the faster setting."""
    best = None
    for th in (os.cpu_count() or 1, 1):
        code = (f"import os,sys;os.environ['OMP_NUM_THREADS']='{th}';sys.path.insert(0,{ROOT!r});"
                "import bench;v,dt,t,_=bench.cpu_cycles(1025,1);print(v)")
        try:
            v = float(subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120).stdout.split()[-1])
        except Exception:
            continue
        if best is None or v > best[1]:
            best = (th, v)
    return best[0] if best else 1


def reference_python_timing():
    """The reference's own Python classes timed in the build container (tools/time_reference_python.py; they cannot
    travel to the GPU box): quoted beside the port so that both ends of "the reference's CPU path" are visible."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "reference_python_timing.json")))
        return {"value": d["value"], "unit": d["unit"], "cores": 1, "where": d["where"],
                "sample": "; ".join(f"{r['n']}x{r['n']}: {r['cycles']} cycles in {r['seconds']:.2f} s" for r in d["runs"])}
    except Exception:
        return None


def run_cpu_baseline(n: int, cycles: int, warmup: int = 1):
    th = pick_cpu_threads()
    code = (f"import os,sys,json;os.environ['OMP_NUM_THREADS']='{th}';sys.path.insert(0,{ROOT!r});"
            f"import bench;v,dt,t,h=bench.cpu_cycles({n},{cycles},warmup={warmup});print(json.dumps([v,dt,t,h]))")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=1800)
    v, dt, t, hist = json.loads(out.stdout.strip().splitlines()[-1])
    return {"value": v, "unit": UNIT, "cores": int(t), "kind": "port",
            "sample": f"{cycles} fp64 V(2,2) red-black GS cycles on {n}x{n} after {warmup} warm-up (C/OpenMP restatement "
                      f"of the reference loops, oracle/mg_oracle.c; {dt:.1f} s; host has {os.cpu_count()} logical CPUs)",
            "seconds": dt, "final_residual": hist[-1], "reference_python": reference_python_timing()}


def reference_grid(a) -> int:
    """Grid of the reference arm: the configured one when the host can hold it (the oracle keeps ~8 fp64 arrays per
    level: ~23 GB at 16385^2), else the largest 2^k+1 below it that fits."""
    if a.cpu_n:
        return a.cpu_n
    n = a.n
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 32 << 30
    while n > 129 and 11.0 * n * n * 8 > 0.6 * avail:
        n = (n - 1) // 2 + 1
    return n


def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cyc = max(1, a.steps)
    n_ref = reference_grid(a)
    wu = max(0, a.warmup)
    base = run_cpu_baseline(n_ref, cyc, warmup=wu)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": a.gpus,
            "steps": cyc, "warmup": wu, "ms_per_step": base["seconds"] / cyc * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"2D Poisson {n_ref}x{n_ref} manufactured sin*sin, V(2,2) red-black GS cycles, fp64 "
                                   f"(the reference's CPU arithmetic: C/OpenMP port of its loops) "
                                   + ("(BASELINE configs[2])" if n_ref == a.n else
                                      f"-- bounded sample of the {a.n}x{a.n} config; unknowns/s is size-independent")},
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample", "reference_python")},
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
def gpu_arm(a):
    import numpy as np
    import torch
    import torch.distributed as dist

    from mixed_precision_multigrid_solvers_for_pdes_b200 import MixedPrecisionMultigrid, PoissonProblem, _lib, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = a.n
    peak, peak_src = load_peaks()
    # the reference's tolerance; where it lies below the fp64 rounding floor of the residual evaluation (16385^2) the
    # solve ends on the floor rule of solvers/policy.py and says so in config.stopped_on
    tol = a.tolerance if a.tolerance is not None else 1e-8

    if world > 1:
        from mixed_precision_multigrid_solvers_for_pdes_b200.distributed import run_distributed_bench
        line = run_distributed_bench(a, world, rank, dev, peak, peak_src, ClockSampler)
        if rank == 0:
            print(json.dumps(line), flush=True)
        import gc
        gc.collect()  # captured graphs hold NCCL work: make sure they are gone before the communicator
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
        return

    from mixed_precision_multigrid_solvers_for_pdes_b200.solvers.policy import CONTINUE
    coefficient = None
    if a.varcoef:
        coefficient = lambda X, Y: 1.0 + 0.5 * np.sin(2 * np.pi * X) * np.cos(np.pi * Y) + X * Y  # noqa: E731
        a.no_e2e = True  # no closed-form solution for this operator: the e2e leg's error check has nothing to check
    solver = MixedPrecisionMultigrid(precision_strategy=a.strategy, switch_threshold=1e-6, tolerance=tol,
                                     cycle_type=a.cycle, loader=a.loader, max_iterations=10 ** 9, device=dev,
                                     smoother="red_black_gauss_seidel" if a.varcoef else a.smoother,
                                     use_cuda_graphs=not a.no_graphs, coefficient=coefficient,
                                     use_fused_defect_down=not a.no_dd)
    solver.setup(n, n)
    eng, g = solver._engine, solver._grid
    b64 = eng.levels[0].bufs(torch.float64)
    ops.fill_sinsin_(b64.f, (0.0, 1.0, 0.0, 1.0), 2 * np.pi ** 2, 1.0, 1.0)
    ops.zero_ring_(b64.f)

    # the solve loop of MixedPrecisionMultigrid.solve, unrolled into steps: same policy object, same launches
    state = {"pol": None, "first": True, "solves": 0, "cycles_per_solve": [], "precisions": [], "stopped_on": None}

    def restart():
        pol = state["pol"] = solver.make_policy()
        state["first"] = True  # the zero initial guess is neither memset nor read (U_ZERO passes)
        if pol.phase == "refine":
            solver._refinement_residual(u_zero=True)
        elif pol.phase == "fp32":
            ops.cast(b64.f, torch.float32, out=eng.levels[0].bufs(torch.float32).f)

    def step():
        pol = state["pol"]
        ph, first = pol.phase, state["first"]
        state["first"] = False
        if ph == "refine":
            norm = solver._cycle_refinement(u_zero=first, last_hint=pol.likely_last())
        elif ph == "fp64":
            norm = solver._cycle_fp64(u_zero=first)
        else:
            norm = solver._cycle_fp32_only(u_zero=first)
        state["precisions"].append(ph)
        if pol.observe(norm) != CONTINUE or len(pol.history) >= 30:
            state["solves"] += 1
            state["cycles_per_solve"].append(len(pol.history))
            state["last_hist"] = list(pol.history)
            state["stopped_on"] = pol.stopped_on
            state["floor_bound"] = pol.floor_bound
            restart()

    restart()
    # setup (untimed, like a compile step): whole solves until every (phase, buffer-role) CUDA graph the solve
    # loop replays has been captured, before the warm-up steps start
    from mixed_precision_multigrid_solvers_for_pdes_b200.solvers.graphs import prime
    primed = prime(step, solver._graph_cache, lambda: state["solves"])
    for _ in range(max(3, a.warmup)):
        step()
    torch.cuda.synchronize()
    launches0 = _lib.call("mg_launch_count")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        torch.cuda.synchronize()
        clk.mark(0)
        e0.record()
        for _ in range(a.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        clk.mark(1)
        ms = e0.elapsed_time(e1)
        # keep the GPU under the same load until the sampler (200 ms period) has seen it for >= 1.5 s; untimed
        t_end = time.perf_counter() + (max(0.0, 1.5 - ms * 1e-3) if clk.thread is None else 0.0)  # nvidia-smi fallback only
        while time.perf_counter() < t_end:
            step()
        torch.cuda.synchronize()
    replayed = solver._graph_cache.captured
    # kernels launched per step: counted by the library on an eager pass of the same steps below
    value = n * n * a.steps / (ms * 1e-3)

    # per-kernel durations: the same K steps once more, eagerly (CUDA events cannot be recorded inside a
    # replayed graph), with an event pair around every level-0 launch on the launching stream
    solver.use_cuda_graphs = False
    ops.TIMER = ops.KernelTimer(min_points=0)  # every level: the cycle's own algorithmic bytes come from this list
    launches0 = _lib.call("mg_launch_count")
    t0e = torch.cuda.Event(enable_timing=True)
    t1e = torch.cuda.Event(enable_timing=True)
    t0e.record()
    for _ in range(a.steps):
        step()
    t1e.record()
    torch.cuda.synchronize()
    ms_eager = t0e.elapsed_time(t1e)
    launches = _lib.call("mg_launch_count") - launches0
    kern = ops.TIMER.summary()
    ops.TIMER = None
    solver.use_cuda_graphs = not a.no_graphs

    # roofline of the dominant kernel (largest total time among the fused passes)
    def alg_bytes(tag):
        """Algorithmic HBM bytes of one launch, from the pass's own grid."""
        name, dt, dims = tag.split("/")
        px, py = (int(v) for v in dims.split("x"))
        pts = px * py
        w = 8 if dt == "f64" else 4
        if name.startswith("var:"):       # variable coefficients: + the nodal coefficient row, once per pass
            return alg_bytes(name[4:] + "/" + dt + "/" + dims) + w * pts
        if name.startswith("small"):      # whole coarse sub-cycle in shared memory: read f (+u), write u
            return 2.0 * w * pts
        if name.startswith("dd:"):        # fused defect + down pass: the defect pass's traffic + e' and f_c out
            return alg_bytes(name[3:].replace("+rbgs2+R", "") + "/" + dt + "/" + dims) + (4 + 1) * pts
        if "resid32" in name or "update" in name:
            b = 0.0
            if "resid32" in name:
                b += 8 + 8 + 4            # read u64, f64; write r32
            if "update" in name:
                b += 4 + 8                # read e32; write u64
            if name.startswith("Z+"):
                b -= 8                    # the zero iterate is not read
            return b * pts
        b = 3.0 * w                       # read u, read f, write u
        if name.startswith("Z+"):
            b -= w                        # the zero iterate is not read
        if "P+" in name:
            b += 0.25 * w                 # read the coarse correction
        if "+R" in name:
            b += 0.25 * w                 # write the restricted residual
        return b * pts
    roof = None
    kernels = {}
    cycle_bytes = 0.0
    for tag, d in sorted(kern.items(), key=lambda kv: -kv[1]["total_ms"]):
        cycle_bytes += alg_bytes(tag) * d["launches"]
        if not tag.endswith(f"/{n}x{n}"):
            continue                      # the per-kernel table lists level 0; coarser levels enter cycle_roofline
        ach = alg_bytes(tag) / (d["mean_ms"] * 1e-3) / 1e9
        sm = "rbgs" if "rbgs" in tag else ("jac" if "jac" in tag else None)
        sweeps = int(tag.split(sm)[1][0]) if sm else 0
        w = 8 if "/f64/" in tag else 4
        kernels[tag] = {"launches": d["launches"], "mean_ms": round(d["mean_ms"], 4), "hbm_gbs": round(ach, 1),
                        "frac_of_peak": round(ach / peak, 4), "share_of_step": round(d["total_ms"] / ms_eager, 4),
                        "smoother_alg_gbs": round(3 * w * sweeps * n * n / (d["mean_ms"] * 1e-3) / 1e9, 1)}
        if roof is None:
            roof = {"bound": "hbm", "kernel": tag, "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                    "frac": round(ach / peak, 4), "traffic": None, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes(tag)}
    coarse_ms = sum(d["total_ms"] for t, d in kern.items() if not t.endswith(f"/{n}x{n}")) / a.steps
    tr = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if roof is not None and os.path.exists(tr):
        try:
            roof["traffic"] = json.load(open(tr)).get(roof["kernel"])
        except Exception:
            pass
    # whole-cycle roofline from the bytes of the passes that actually ran (all levels, restarts included)
    bpu = cycle_bytes / a.steps / (n * n)
    cycle_frac = bpu * value / 1e9 / peak
    assert cycle_frac <= 1.05, f"cycle roofline fraction {cycle_frac:.3f} > 1: the byte model is wrong"

    # end to end through the public API with HOST buffers (pinned): H2D of f, solve, D2H of u
    e2e = None
    if not a.no_e2e:
        f_host = torch.empty((n, n), dtype=torch.float64, pin_memory=True)
        f_host.copy_(b64.f)
        torch.cuda.synchronize()
        api = MixedPrecisionMultigrid(precision_strategy=a.strategy, switch_threshold=1e-6, tolerance=tol,
                                      cycle_type=a.cycle, loader=a.loader, device=dev, use_cuda_graphs=not a.no_graphs,
                                      use_fused_defect_down=not a.no_dd)
        api._engine, api._shape, api._domain, api._grid, api._sumsq, api._pinned_out, api._graph_cache = (
            solver._engine, solver._shape, solver._domain, solver._grid, solver._sumsq, None, solver._graph_cache)
        prob = PoissonProblem(rhs=f_host, nx=n, ny=n)
        for _ in range(3):  # warm-up: pinned result staging; a CUDA graph replays from its third use (eager, capture, replay)
            api.solve(prob)
        reps, tot_t, tot_c, info, each = 3, 0.0, 0, None, []
        u_host = None
        for _ in range(reps):
            u_host = None  # the previous result is the caller's to drop: one owned array alive at a time
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            u_host, info = api.solve(prob)
            each.append(time.perf_counter() - t0)
            tot_t += each[-1]
            tot_c += info["iterations"]
        single = {"value": n * n * tot_c / tot_t, "seconds_per_solve": tot_t / reps,
                  "seconds_each": [round(t, 4) for t in each]}
        u_host = None
        # the batch API: B solves, each with its own H2D of f and D2H of u inside the timed region; transfers of
        # neighbouring solves overlap the cycles (copy engines, side streams)
        B = 10  # fill and drain of the pipeline (one un-overlapped upload, one download) amortised over the batch
        outs = [torch.empty((n, n), dtype=torch.float64, pin_memory=True) for _ in range(2)]
        probs = [prob] * B
        api.solve_many(probs[:2], outputs=outs)  # warm-up: staging buffers, streams
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, infos = api.solve_many(probs, outputs=[outs[k % 2] for k in range(B)])
        torch.cuda.synchronize()
        tb = time.perf_counter() - t0
        cyc = sum(i["iterations"] for i in infos)
        info = infos[-1]
        e2e = {"value": n * n * cyc / tb, "unit": UNIT, "h2d_bytes_per_step": n * n * 8,
               "d2h_bytes_per_step": n * n * 8 + 8 * info["iterations"],
               "step": "one solve of a solve_many() batch of %d: pinned host f -> device, %d cycles, device u -> pinned "
                       "host; the transfers of neighbouring solves overlap the cycles" % (B, info["iterations"]),
               "seconds_per_solve": tb / B, "iterations": info["iterations"], "final_residual": info["final_residual"],
               "max_error": float(ops.maxerr_sinsin(eng.levels[0].bufs(torch.float64).u)),
               "single_solve": {"value": single["value"], "seconds_per_solve": single["seconds_per_solve"],
                                "seconds_each": single["seconds_each"],
                                "step": "one solve() call, nothing overlapped: H2D of f, cycles, D2H of u back to back"}}
        del outs
        del f_host

    cpu = None if a.no_cpu_baseline else run_cpu_baseline(a.cpu_n or 8193, 8)
    clocks = clk.summary()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": a.steps, "warmup": max(3, a.warmup),
        "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 cycle / f64 iterate+residual" if solver.mode in ("switch", "refine") else
                 ("f64" if solver.mode == "fp64" else "f32"),
        "data": "synthetic",
        "config": {"workload": ("2D variable-coefficient -div(a grad u) = f " if a.varcoef else "2D Poisson ") +
                               f"{n}x{n} manufactured sin*sin, {a.cycle}(2,2) "
                               f"{'red-black GS' if a.smoother == 'rbgs' else 'damped Jacobi (2/3)'}, "
                               f"precision_strategy={a.strategy} (BASELINE configs[2])", "levels": eng.num_levels,
                   "loader": a.loader, "cuda_graphs": (not a.no_graphs), "graphs_captured": replayed, "priming_solves": primed,
                   "kernel_timing": "eager replay of the same %d steps with CUDA events around each level-0 launch "
                                    "(%.3f ms/step eager)" % (a.steps, ms_eager / a.steps), "tolerance": tol, "switch_threshold": 1e-6, "l2": "inputs (>= 1 GB per array) exceed the 126 MB L2; no flush needed",
                   "cycles_per_solve": state["cycles_per_solve"][-3:], "last_residual_history": state.get("last_hist"),
                   "stopped_on": state.get("stopped_on"), "attainable_residual_bound": state.get("floor_bound")},
        "roofline": roof, "kernels": kernels,
        "cycle_roofline": {"bytes_per_unknown": round(bpu, 2), "frac_of_peak": round(cycle_frac, 4),
                           "note": "algorithmic bytes of every pass launched per step (all levels, restart passes "
                                   "included) / unknowns; whole-step time from the graph-replayed timed region",
                           "coarse_levels_ms_per_step_eager": round(coarse_ms, 4),
                           "survey_fp64_model_bytes_per_unknown": BYTES_PER_UNKNOWN_FP64_V22},
        "cpu_baseline": ({k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "reference_python")} if cpu else None),
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }
    print(json.dumps(line), flush=True)


def main():
    a = parse()
    if a.impl == "reference":
        reference_arm(a)
    else:
        gpu_arm(a)


if __name__ == "__main__":
    main()
