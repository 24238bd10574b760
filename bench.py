#!/usr/bin/env python
"""bench.py — reactor ODE solves/s on the BASELINE config-3 workload.

Workload (BASELINE.json configs[2], SURVEY.md 8(d) "config 3"): an ensemble of
steady-state pseudo-homogeneous packed-bed reactors (PyREMOT model N1,
CO2-to-methanol/DME, 6 species / 3 reactions / 8 unknowns) sweeping feed
temperature, pressure and composition; 2^20 reactors PER GPU (weak scaling),
SciPy-default tolerances rtol=1e-3 / atol=1e-6 — what the reference's
`solve_ivp` call uses (pbHomoReactor.py:2931) — outlet-only output.

A "step" is one pass of the hot path over the batch: the per-reactor setup
kernel (runN1 :2744-2852) + the fused adaptive-Rosenbrock integrator kernel
(solve_ivp + modelEquationN1 + sortResult4).  `value` is timed with the inputs
resident in HBM; `e2e` goes through the public `rmtExeBatch` call with host
arrays (pinned staging + H2D + kernels + D2H inside the timed region).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402  (synthetic inputs shared with the tests)

METRIC = "reactor_ode_solves_per_sec"
UNIT = "solves/s"
B_PER_GPU = 1 << 20
SEED = 20240611
RTOL, ATOL = 1e-3, 1e-6


# ------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation of the path on the host cores.
# kind "reference" = the UNMODIFIED PyREMOT package staged under oracle/_ref by oracle/stage_reference.py (build());
# kind "port"      = oracle/pyremot_oracle.py (same equations, same SciPy call, without the reference's per-call
#                    `eval` of Cp strings — about 8x faster per solve), used when nothing is staged and reported
#                    beside the reference figure otherwise.
# ------------------------------------------------------------------------------------
def _cpu_paths():
    for p in (os.path.join(ROOT, "oracle"),):
        if p not in sys.path:
            sys.path.insert(0, p)


def reference_staged():
    _cpu_paths()
    import ref_harness
    return ref_harness.available()


def _case_inputs(case, n):
    """(base modelInput, sweep) of the CPU sample: the first n instances of the GPU arm's own draw."""
    if case == "config4":
        base = cases.methanol_readme_input("N1")
        base["reaction-rates"] = cases.methanol_kinetics_param(1171.2)
        return base, cases.config4_population(n)
    return cases.methanol_readme_input("N1"), cases.config3_sweep(n, SEED)


def _cpu_worker(job):
    kind, case, idx_chunk = job
    _cpu_paths()
    import io
    import contextlib
    import warnings
    warnings.simplefilter("ignore")
    base, sweep = _case_inputs(case, max(idx_chunk) + 1)
    nfev, ok = 0, 0
    if kind == "reference":
        import ref_harness as R
        for i in idx_chunk:
            try:
                with R.Capture() as cap:
                    R.rmtExe(cases.instance_input(base, sweep, i))
                nfev += cap.calls[0]["nfev"]
                ok += 1
            except Exception:
                pass
        return nfev, ok
    import pyremot_oracle as O
    for i in idx_chunk:
        mi = cases.instance_input(base, sweep, i)
        with contextlib.redirect_stdout(io.StringIO()):
            try:
                r = O.rmtExe(mi)
                nfev += r["resModel"][0]["nfev"]
                ok += 1
            except Exception:
                pass
    return nfev, ok


def cpu_solves(n_solves, pool, cores, kind="port", case="config3"):
    chunks = [list(range(c, n_solves, cores)) for c in range(cores)]
    jobs = [(kind, case, c) for c in chunks if c]
    t0 = time.perf_counter()
    res = pool.map(_cpu_worker, jobs)
    dt = time.perf_counter() - t0
    return dt, sum(r[0] for r in res), sum(r[1] for r in res)


def cpu_baseline_block(cores, pool, per_core_ref=6, per_core_port=16):
    """`cpu_baseline` of the GPU arm's line: the reference (when staged) on a bounded sample of the config-3 draw,
    the port's figure beside it."""
    kind = "reference" if reference_staged() else "port"
    out = {}
    cpu_solves(cores, pool, cores, "port")                        # warm the workers (imports)
    S = per_core_port*cores
    dt, nfev, ok = cpu_solves(S, pool, cores, "port")
    port = {"value": S/dt, "unit": UNIT, "cores": cores, "sample": "%d config-3 reactors" % S,
            "mean_nfev": nfev/max(ok, 1), "converged": ok}
    if kind == "reference":
        cpu_solves(cores, pool, cores, "reference")
        S = per_core_ref*cores
        dt, nfev, ok = cpu_solves(S, pool, cores, "reference")
        out = {"value": S/dt, "unit": UNIT, "cores": cores, "kind": "reference",
               "sample": "%d config-3 reactors (first indices of the GPU arm's own seed-%d draw, identical float64 inputs), the "
                         "UNMODIFIED PyREMOT rmtExe (oracle/_ref, staged byte for byte by oracle/stage_reference.py; matplotlib "
                         "stubbed, display-result False, stdout discarded), SciPy LSODA at its default rtol=1e-3 atol=1e-6, "
                         "one process per core over %d cores; mean nfev %.0f; %d/%d converged"
                         % (S, SEED, cores, nfev/max(ok, 1), ok, S),
               "solves_per_s_per_core": S/dt/cores, "port": port}
    else:
        out = dict(port, kind="port")
        out["sample"] += (" (first indices of the same seed), oracle port of runN1/modelEquationN1 + SciPy LSODA rtol=1e-3 "
                          "atol=1e-6, %d processes; the reference is not staged under oracle/_ref" % cores)
    return out


def single_reference_times():
    """Config 1 (one N1 solve through the reference's rmtExe, best of 3) and the reference's RHS alone
    (modelEquationN1 called on states of its own solution) on ONE host core."""
    _cpu_paths()
    import ref_harness as R
    mi = cases.methanol_readme_input("N1")
    best, nfev = None, 0
    for _ in range(3):
        with R.Capture(keep_fun=True) as cap:
            _, w = R.rmtExe(mi)
        best = w if best is None else min(best, w)
        c = cap.calls[0]
        nfev = c["nfev"]
    fun, ps, Y = c["fun"], c["args"][0], c["y"]
    import io
    import contextlib
    n = 0
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        while time.perf_counter() - t0 < 2.0:
            for k in range(Y.shape[1]):
                fun(0.0, list(Y[:, k]), ps)
            n += Y.shape[1]
    rhs_s = (time.perf_counter() - t0)/n
    return {"config1_single_rmtExe_s": best, "config1_nfev": int(nfev), "modelEquationN1_s_per_call": rhs_s,
            "modelEquationN1_evals_per_s_per_core": 1.0/rhs_s, "rhs_calls_timed": n}


def run_reference_arm(args, rank, world):
    """`--impl reference`: the reference's own CPU implementation of the path (the unmodified package staged under
    oracle/_ref; the oracle port when nothing is staged) on all host cores, default tolerances, on a bounded sample
    of the GPU arm's workload per step."""
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    kind = "reference" if reference_staged() else "port"
    per_step = (2 if kind == "reference" else 8)*cores
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        cpu_solves(cores, pool, cores, kind)                      # imports
        for _ in range(args.warmup):
            cpu_solves(cores, pool, cores, kind)
        t_total, nfev, ok = 0.0, 0, 0
        for _ in range(args.steps):
            dt, nf, k = cpu_solves(per_step, pool, cores, kind)
            t_total += dt; nfev += nf; ok += k
    value = args.steps*per_step/t_total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3*t_total/args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": "%d config-3 reactors per step (first indices of seed %d; each warm-up step %d), %s, SciPy LSODA "
                                   "rtol=1e-3 atol=1e-6, multiprocessing over %d cores; mean nfev %.0f; %d/%d converged"
                                   % (per_step, SEED, cores,
                                      "UNMODIFIED PyREMOT rmtExe from oracle/_ref" if kind == "reference" else "oracle port",
                                      cores, nfev/max(ok, 1), ok, args.steps*per_step)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def run_cpu_baselines(args):
    """`--cpu-baselines`: the long CPU measurements of BASELINE.md section 3 (minutes; not part of the default run).
    Prints one JSON object; the copy measured on the GPU box's host is committed as profiles/r02_cpu_baselines.json and
    quoted (with that provenance) by the default line."""
    import multiprocessing as mp
    import platform
    _cpu_paths()
    import ref_harness as R
    cores = os.cpu_count() or 1
    out = {"host": {"cores": cores, "machine": platform.machine(), "python": platform.python_version()},
           "reference": "unmodified PyREMOT from " + str(R.reference_root())}
    try:
        out["host"]["model_name"] = [l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except Exception:
        pass
    out.update(single_reference_times())
    with mp.get_context("spawn").Pool(cores) as pool:
        cpu_solves(cores, pool, cores, "reference")
        for case, S in (("config3", 512), ("config4", 256)):
            dt, nfev, ok = cpu_solves(S, pool, cores, "reference", case)
            out[case] = {"sample": S, "seconds": dt, "solves_per_s_box": S/dt, "solves_per_s_per_core": S/dt/cores,
                         "mean_nfev": nfev/max(ok, 1), "converged": ok,
                         "extrapolated_full_size_hours": ((1 << 20) if case == "config3" else 65536)/(S/dt)/3600.0}
        dt, nfev, ok = cpu_solves(512, pool, cores, "port", "config3")
        out["config3_port"] = {"sample": 512, "seconds": dt, "solves_per_s_box": 512/dt, "mean_nfev": nfev/max(ok, 1)}
    if not args.no_n2:
        # config 2: one N2 instance, 50 axial nodes, README inputs, BDF (LSODA needs longer)
        mi2 = cases.methanol_readme_input("N2")
        mi2["solver-config"]["ivp"] = "BDF"
        R.set_grid("N2", zNo=50)
        with R.Capture() as cap:
            _, w = R.rmtExe(mi2)
        R.set_grid("N2", zNo=20)
        out["config2_single_n2_50_nodes_bdf_s"] = w
        out["config2_nfev"] = int(sum(c["nfev"] for c in cap.calls))
        out["config2_njev"] = int(sum(c["njev"] for c in cap.calls))
        # config 5: the reference at 200 nodes needs > 1 h per instance (FD Jacobian of 1400 unknowns = 1400 RHS calls
        # of ~0.25 s); quote its RHS cost at 200 nodes and the 50-node run instead
        R.set_grid("N2", zNo=200)
        mi5 = cases.instance_input(cases.methanol_readme_input("N2"), cases.config3_sweep(8, 20240613), 0)
        try:
            with R.Capture(capture_only=True) as cap:
                try:
                    R.rmtExe(mi5)
                except R.StopAfter:
                    pass
            c = cap.calls[0]
            import io
            import contextlib
            t0 = time.perf_counter(); k = 0
            with contextlib.redirect_stdout(io.StringIO()):
                while time.perf_counter() - t0 < 3.0:
                    c["fun"](0.0, list(c["y0"]), c["args"][0]); k += 1
            out["config5_modelEquationN2_200_nodes_s_per_call"] = (time.perf_counter() - t0)/k
        except Exception as e:                                       # diagnostics only
            out["config5_rhs_error"] = repr(e)
        R.set_grid("N2", zNo=20)
    emit(out)


def workload_config(world):
    cfg = {
        "workload": "config3: 2^20 steady-state PFR (PyREMOT N1, CO2->MeOH/DME, 6 comps, 3 rxns, 8 unknowns) per GPU; "
                    "T0~U[473,573]K, P0~U[2,8]MPa, H2/COx~U[1,3], CO2/COx~U[0.2,0.8]; seed %d+rank" % SEED,
        "instances_per_gpu": B_PER_GPU, "rtol": RTOL, "atol": ATOL, "output": "outlet (y_i, P, T)",
        "integrator": "auto -> Ros4(3) L-stable Rosenbrock (outlet only, rtol >= 5e-4; Rodas4(3) otherwise), analytic Jacobian", "parallelism": "ensemble-sharded x%d, no data-path collective" % world,
        "cache": "inputs+constants+outputs 410 MB per step > 126 MB L2 (no flush needed)",
    }
    return cfg


# ------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def mark(self, wait_s=4.0):
        """Call right before the timed region: waits for the sampler's first line and remembers
        how many lines precede the region."""
        self.n0 = 0
        if self.proc is None:
            return
        t0 = time.time()
        while time.time() - t0 < wait_s:
            try:
                n = sum(1 for _ in open(self.path))
            except Exception:
                n = 0
            if n > 0:
                self.n0 = n
                return
            time.sleep(0.05)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons, mx = [], set(), None
        try:
            for ln in list(open(self.path))[getattr(self, "n0", 0):]:
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx = float(f[2])
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------
CALIBRATION = os.path.join(ROOT, "profiles", "r02_calibration.json")
CPU_BASELINES = os.path.join(ROOT, "profiles", "r02_cpu_baselines.json")


def load_json(path):
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return None


def calibration_for(cal, kernel, module_key):
    """The ncu-derived numbers of one kernel (tools/make_calibration.py wrote them from profiles/*.csv together with
    the identity of the cubin they were measured on).  Returns (entry or None, stale flag)."""
    if not cal or kernel not in cal.get("kernels", {}):
        return None, True
    e = cal["kernels"][kernel]
    return e, e.get("module_key") != module_key


def solver_flops(info, stats, cm):
    """Algorithmic flops of one integrator launch from its per-instance counters.
    Per attempt: 1 evaluation of g and A = dg/dx, the other RHS evaluations, one m x m LU,
    s triangular solves, stage combinations and the error norm; m = nr + 2 unknowns when the
    integrator works in reaction extents, else n.  (alg: 1 flop per op; wt: FP64-instruction
    weighted, rmt_app_b200/expr.py FLOP_WEIGHT.)"""
    att = float(stats[3].sum())
    nfev = float(stats[2].sum())
    s = info.stages
    n = cm.m
    lu = (2.0*n**3)/3.0 + n*n            # factorisation incl. forming W
    tri = s*2.0*n*n
    comb = 2.0*n*(s*(s - 1)) + 8.0*info.n  # a_ij / c_ij combinations, update, error norm (upper bound: all coefficients)
    if cm.reduced:
        comb += (s + 1)*2.0*float((cm.spec.nu != 0).sum())      # E x: stage arguments, y_new and the error vector
    lin_alg = att*(lu + tri + comb)
    lin_wt = lin_alg + att*9.0*n          # n reciprocal pivots
    alg = nfev*info.flops_rhs_alg + att*cm.flops["sys_alg"] + lin_alg
    wt = nfev*info.flops_rhs_wt + att*cm.flops["sys_weighted"] + lin_wt
    return alg, wt, att, nfev


def run_gpu_arm(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from rmt_app_b200 import capi, engine, ensemble, rmtExe, rmtExeBatch, solverSetting

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(vals):
        t = torch.tensor([float(v) for v in vals], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]

    cal = load_json(CALIBRATION)
    cpu_file = load_json(CPU_BASELINES)

    # CPU baseline first (rank 0, N=1 only), before the timed GPU region
    cpu_baseline = None
    cpu_single = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import multiprocessing as mp
        cores = os.cpu_count() or 1
        with mp.get_context("spawn").Pool(cores) as pool:
            cpu_baseline = cpu_baseline_block(cores, pool)
        if reference_staged():
            cpu_single = single_reference_times()

    B = args.instances
    base = cases.methanol_readme_input("N1")
    sweep = cases.config3_sweep(B, SEED + rank)
    # outlet-only at rtol 1e-3: the automatic choice is the 4-stage Ros4 tableau (engine.choose_method)
    cm = engine.compile_model(base, method=engine.choose_method(base, RTOL, 1))
    ctrl = engine.METHOD_CTRL.get(cm.method)
    mod = cm.load(local_rank)
    info, spec = mod.info, cm.spec
    n = info.n
    stream = torch.cuda.current_stream().cuda_stream
    ws = engine.Workspace()

    # resident inputs
    h_rows, n_rows, row_map = engine.sweep_rows_into(spec, sweep, B, ws)
    uniform = engine.uniform_inputs(spec, base)
    d_rows = h_rows.to(dev)
    d_consts = torch.empty((info.nconst, B), dtype=torch.float64, device=dev)
    d_out = torch.empty((1, n, B), dtype=torch.float64, device=dev)
    d_status = torch.empty((B,), dtype=torch.int32, device=dev)
    d_stats = torch.empty((4, B), dtype=torch.int32, device=dev)
    z_eval = np.array([1.0])

    def step(ev=None):
        if ev is not None:
            ev[0].record()
        mod.setup(B, d_rows, n_rows, row_map, uniform, d_consts, stream=stream)
        if ev is not None:
            ev[1].record()
        mod.n1_solve(B, d_consts, z_eval, RTOL, ATOL, d_out, d_status, d_stats, out_mode=1, ctrl=ctrl, stream=stream)
        if ev is not None:
            ev[2].record()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    sampler.mark()
    barrier()
    t_start = torch.cuda.Event(enable_timing=True); t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    for k in range(args.steps):
        step(evs[k])
    t_end.record()
    barrier()
    clocks = sampler.stop()
    elapsed_ms = max_over_ranks(t_start.elapsed_time(t_end))
    setup_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in evs]))
    solve_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in evs]))

    status = d_status.cpu().numpy()
    stats = d_stats.cpu().numpy()
    n_ok = int((status == 0).sum())
    alg, wt, att, nfev = solver_flops(info, stats, cm)
    n_ok_all = int(sum_over_ranks([n_ok])[0])

    # ---- e2e through the public API: host arrays in, host arrays out -----------------------------
    # inputs live in pinned host memory (the contract's e2e definition); results land in pinned host memory
    psweep = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in sweep.items()}
    for _ in range(2):
        r = rmtExeBatch(base, psweep, workspace=ws, rtol=RTOL, atol=ATOL, return_stats=False)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = rmtExeBatch(base, psweep, workspace=ws, rtol=RTOL, atol=ATOL, return_stats=False)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    h2d, d2h = r["h2d_bytes"], r["d2h_bytes"]
    e2e_ok = int(r["success"].sum())

    # ---- the same through the C ABI's host-buffer entry point (rmt_n1_solve_host: no torch on the path) -----------
    hp_out = torch.empty((1, n, B), dtype=torch.float64).pin_memory()
    hp_status = torch.empty((B,), dtype=torch.int32).pin_memory()
    rows_np, out_np, status_np = h_rows.numpy(), hp_out.numpy(), hp_status.numpy()
    for _ in range(2):
        mod.n1_solve_host(B, rows_np, n_rows, row_map, uniform, z_eval, RTOL, ATOL, out_np, status_np, ctrl=ctrl)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        mod.n1_solve_host(B, rows_np, n_rows, row_map, uniform, z_eval, RTOL, ATOL, out_np, status_np, ctrl=ctrl)
    cabi_s = max_over_ranks(time.perf_counter() - t0)
    cabi_ok = int((status_np == 0).sum())

    # ---- the parity-grade tolerance (what the 1e-6 agreement with the reference is proven at) ----------------------
    tight = None
    if not args.no_tight:
        cmt = engine.compile_model(base, method=engine.choose_method(base, 1e-9, 1))          # Rodas4(3)
        modt = cmt.load(local_rank)
        ctrl_t = engine.METHOD_CTRL.get(cmt.method)
        Bt = B
        for _ in range(2):
            modt.n1_solve(Bt, d_consts, z_eval, 1e-9, 1e-12, d_out, d_status, d_stats, out_mode=1, ctrl=ctrl_t, stream=stream)
        barrier()
        q0 = torch.cuda.Event(enable_timing=True); q1 = torch.cuda.Event(enable_timing=True)
        q0.record()
        modt.n1_solve(Bt, d_consts, z_eval, 1e-9, 1e-12, d_out, d_status, d_stats, out_mode=1, ctrl=ctrl_t, stream=stream)
        q1.record(); torch.cuda.synchronize()
        tms = max_over_ranks(q0.elapsed_time(q1))
        st_t = d_stats[:, :Bt].cpu().numpy()
        ok_t, att_t = sum_over_ranks([int((d_status[:Bt] == 0).sum().item()), float(st_t[3].sum())])
        tight = {"instances": world*Bt, "rtol": 1e-9, "atol": 1e-12, "integrator": cmt.method + " (6 stages, stiffly accurate)",
                 "ms": tms, "solves_per_s": world*Bt/(tms*1e-3), "attempts_per_solve": att_t/(world*Bt), "converged": int(ok_t),
                 "note": "the tolerance at which outlets agree with the converged reference to <= 1e-6 "
                         "(tests/test_gpu_n1.py level 2); the same 2^20 config-3 reactors, inputs resident, outlet only"}

    # ---- BASELINE configs[0]: one N1 reactor through rmtExe ---------------------------------------------------------
    config1 = None
    if rank == 0:
        rmtExe(base)
        best = min(_timed(lambda: rmtExe(base)) for _ in range(5))
        config1 = {"rmtExe_single_N1_s": best, "what": "rmt_app_b200.rmtExe(modelInput): README inputs, 101-point profiles, default tolerances, "
                                                       "warm module (trace + NVRTC + module load happen once per process)"}
        if cpu_single:
            config1.update(reference_rmtExe_s=cpu_single["config1_single_rmtExe_s"], reference_nfev=cpu_single["config1_nfev"],
                           speedup=cpu_single["config1_single_rmtExe_s"]/best, reference="unmodified PyREMOT (oracle/_ref), this host, one core")
        elif cpu_file:
            config1.update(reference_rmtExe_s=cpu_file.get("config1_single_rmtExe_s"), reference="profiles/r02_cpu_baselines.json (measured on a GPU box host)")

    # ---- BASELINE configs[3]: parameter-estimation population, sharded, ONE packed collective ----------------------
    config4 = None
    if not args.no_config4:
        base4 = cases.methanol_readme_input("N1")
        base4["reaction-rates"] = cases.methanol_kinetics_param(1171.2)
        B4 = 65536
        pop = cases.config4_population(B4)
        ppop = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in pop.items()}
        cm4 = engine.compile_model(base4, method=engine.choose_method(base4, RTOL, 1))
        nominal = engine.n1_solve_ensemble(cm4, base4, None, 1, rtol=1e-9, atol=1e-12).out[0, :, 0]
        ws4 = engine.Workspace()
        for _ in range(2):
            r4 = ensemble.rmtExeBatchSharded(base4, ppop, B4, objective_ref=nominal, workspace=ws4)
        reps = 10
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            r4 = ensemble.rmtExeBatchSharded(base4, ppop, B4, objective_ref=nominal, workspace=ws4)
        torch.cuda.synchronize()
        ms4 = max_over_ranks((time.perf_counter() - t0)/reps*1e3)
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            r4b = ensemble.rmtExeBatchSharded(base4, ppop, B4, objective_ref=nominal, workspace=ws4, gather=False)
        torch.cuda.synchronize()
        ms4b = max_over_ranks((time.perf_counter() - t0)/reps*1e3)
        config4 = {"parameter_sets": B4, "ms_per_population": ms4, "parameter_sets_per_s": B4/(ms4*1e-3),
                   "ms_per_population_statistics_only": ms4b,
                   "includes": "wall time of ensemble.rmtExeBatchSharded (max over ranks): H2D of the rank's shard from pinned memory, setup, "
                               "solve with fused objective and fused (sum, min, argmin, failed) reduction, "
                               + ("ONE NCCL all-gather of the packed [outlets | objectives | status | tail] buffer" if world > 1
                                  else "no collective (one rank)") + ", one D2H of the gathered buffer; 'statistics_only' gathers the 4-number tails only",
                   "objective_min": r4["objective_min"], "objective_argmin": r4["objective_argmin"], "failed": r4["failed"],
                   "statistics_only_same_minimum": bool(r4b["objective_min"] == r4["objective_min"] and r4b["objective_argmin"] == r4["objective_argmin"])}
        if cpu_file and "config4" in cpu_file:
            c4 = cpu_file["config4"]
            config4["reference"] = {"solves_per_s_box": c4["solves_per_s_box"], "cores": cpu_file["host"]["cores"], "sample": c4["sample"],
                                    "ms_per_population_extrapolated": B4/c4["solves_per_s_box"]*1e3,
                                    "source": "profiles/r02_cpu_baselines.json (unmodified reference, GPU box host)"}

    # ---- strong scaling: ONE fixed ensemble over all ranks, gather to rank 0 included ------------------------------
    strong = None
    if not args.no_strong:
        Bs = args.strong_instances
        sw_s = cases.config3_sweep(Bs, SEED + 1000)                    # the same draw on every rank
        psw_s = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in sw_s.items()}
        wss = engine.Workspace()
        rs = ensemble.rmtExeBatchSharded(base, psw_s, Bs, rtol=RTOL, atol=ATOL, gather="root", workspace=wss)   # warm-up
        barrier()
        t0 = time.perf_counter()
        rs = ensemble.rmtExeBatchSharded(base, psw_s, Bs, rtol=RTOL, atol=ATOL, gather="root", workspace=wss)
        torch.cuda.synchronize()
        ss = max_over_ranks(time.perf_counter() - t0)
        strong = {"instances_total": Bs, "seconds": ss, "solves_per_s": Bs/ss, "failed": rs["failed"],
                  "includes": "ensemble.rmtExeBatchSharded(gather='root'), wall time, max over ranks: every rank copies its contiguous "
                              "1/N of the pinned inputs H2D, solves it, ONE all-gather of outlets + status, rank 0 copies the full "
                              "result (%.0f MB) to the host" % (Bs*(8*n + 4)/1e6),
                  "scaling": "strong (total work fixed as N grows)"}
        del psw_s, sw_s, wss, rs

    # ---- the reference's own output shape: 101-point profiles of every reactor (runN1's t_eval, :2931) -----------
    profile = None
    if rank == 0 and not args.no_profile:
        Bp = min(B, 1 << 18)
        zp = np.linspace(0, 1, 101)
        cmp_ = engine.compile_model(base, method=engine.choose_method(base, RTOL, zp.size))
        dsw = {k: torch.from_numpy(np.ascontiguousarray(v[:Bp])).to(dev) for k, v in sweep.items()}
        wsp = engine.Workspace()
        for _ in range(2):
            rp = engine.n1_solve_ensemble(cmp_, base, dsw, Bp, z_eval=zp, rtol=RTOL, atol=ATOL, keep_on_device=True, workspace=wsp)
        p0 = torch.cuda.Event(enable_timing=True); p1 = torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(3):
            rp = engine.n1_solve_ensemble(cmp_, base, dsw, Bp, z_eval=zp, rtol=RTOL, atol=ATOL, keep_on_device=True, workspace=wsp)
        p1.record(); torch.cuda.synchronize()
        pms = p0.elapsed_time(p1)/3
        profile = {"instances": Bp, "points": int(zp.size), "ms": pms, "solves_per_s": Bp/(pms*1e-3),
                   "output_bytes": int(8*n*zp.size*Bp), "integrator": cmp_.method + ", dense output",
                   "converged": int((rp.status == 0).sum().item()),
                   "note": "what rmtExe returns per reactor (dataYs 8 x 101); inputs and results device-resident"}
        del rp, wsp, dsw

    # ---- roofline denominators measured on this box ----------------------------------------------
    fp64_peak = mod.fp64_peak(iters=16384, repeats=5)
    peaks = load_json(os.path.join(ROOT, "MEASURED_PEAKS.json")) or {}
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md; of fallback)"

    # stand-alone RHS / Jacobian kernels (HBM-bound): L2 flushed before every launch
    d_y = torch.rand((n, B), dtype=torch.float64, device=dev)*0.5 + 0.25
    d_y[n - 2] = 1.0
    d_y[n - 1] = 0.1
    d_f = torch.empty((n, B), dtype=torch.float64, device=dev)
    d_J = torch.empty((n*n, B), dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed_flushed(fn, reps=10):
        ts = []
        for _ in range(reps + 3):
            flush.zero_()                                           # 256 MB > 126 MB L2
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.mean(ts[3:]))
    rhs_ms = timed_flushed(lambda: mod.n1_rhs(B, d_consts, d_y, d_f, stream=stream))
    jac_ms = timed_flushed(lambda: mod.n1_jac(B, d_consts, d_y, d_f, d_J, stream=stream))
    # algorithmic bytes per evaluation: the 14 hot constants + kinetic parameters + state in, derivative (+ Jacobian) out
    rhs_bytes = 8.0*(14 + info.nkp + 2*n)*B
    jac_bytes = 8.0*(14 + info.nkp + 2*n + n*n)*B
    del flush

    # ---- the dynamic model (BASELINE configs[1] and configs[4]) through the product API -----------------------------
    n2 = None
    if not args.no_n2:
        mi2 = cases.methanol_readme_input("N2")
        single_s = None
        if rank == 0:
            solverSetting["N2"]["zNo"] = 50
            rmtExe(mi2)
            single_s = min(_timed(lambda: rmtExe(mi2)) for _ in range(3))
            solverSetting["N2"]["zNo"] = 20
        Bn, zn = 12500*world, 200
        sw2 = cases.config3_sweep(Bn, 20240613)                        # the SAME 100 000-reactor draw on every rank (N = 8)
        psw2 = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in sw2.items()}
        ws2 = engine.Workspace()
        r2 = ensemble.rmtExeBatchN2Sharded(mi2, psw2, Bn, zNo=zn, tNo=5, gather=None, workspace=ws2)        # warm-up
        barrier()
        q0 = torch.cuda.Event(enable_timing=True); q1 = torch.cuda.Event(enable_timing=True)
        q0.record()
        r2 = ensemble.rmtExeBatchN2Sharded(mi2, psw2, Bn, zNo=zn, tNo=5, gather=None, workspace=ws2)
        q1.record(); torch.cuda.synchronize()
        ens_s = max_over_ranks(q0.elapsed_time(q1)*1e-3)
        barrier()
        t0 = time.perf_counter()
        r2g = ensemble.rmtExeBatchN2Sharded(mi2, psw2, Bn, zNo=zn, tNo=5, gather="final", workspace=ws2, keep_on_device=True)
        torch.cuda.synchronize()
        ens_gather_s = max_over_ranks(time.perf_counter() - t0)
        lanes2 = engine.n2_lanes(12500, zn)
        cm2 = engine.compile_model_n2(mi2, 12500, zn)
        i2 = cm2.load(local_rank).info
        att2, ok2, acc2 = sum_over_ranks([float(r2["local_stats"][3].double().sum().item()),
                                          float((r2["local_status"] == 0).sum().item()),
                                          float(r2["local_stats"][0].double().sum().item())])
        nn = i2.n
        # per attempt and node: f + Jacobian blocks, (s-1) RHS, LU + n unit-vector solves (explicit inverse) + the pressure row,
        # per stage one (n+1) x n matrix-vector product and the stage combinations
        lin = (2.0*nn**3)/3.0 + nn*2.0*nn*nn + 2.0*nn*nn + i2.stages*(2.0*(nn + 1)*nn + 4.0*nn*i2.stages)
        alg2 = att2*zn*(i2.flops_jac_alg + (i2.stages - 1)*i2.flops_rhs_alg + lin)
        e2, stale2 = calibration_for(cal, "rmt_n2_solve", cm2.key())
        n2 = {"config2_single_50_nodes_s": single_s,
              "config2_reference_bdf_s": (cpu_file or {}).get("config2_single_n2_50_nodes_bdf_s"),
              "config2_reference_source": "profiles/r02_cpu_baselines.json: the unmodified reference, same inputs, solverSetting zNo = 50, "
                                          "ivp = BDF, measured on a GPU box host (one core)" if cpu_file else None,
              "ensemble": {"instances": Bn, "instances_per_gpu": Bn//world, "nodes": zn, "period_s": 0.5, "seconds": ens_s,
                           "instances_per_s": Bn/ens_s, "converged": int(ok2), "failed": r2["failed"], "steps_mean": acc2/Bn,
                           "lanes_per_reactor": lanes2, "block": cm2.block,
                           "seconds_incl_gather_of_final_profiles": ens_gather_s,
                           "gathered_bytes_per_rank": int(8*r2g["layout"].length*world),
                           "node_rhs_evals_per_s": att2*zn*i2.stages/ens_s,
                           "fp64_tflops_algorithmic": alg2/ens_s/1e12, "fp64_frac_of_measured_peak": alg2/ens_s/1e12/fp64_peak/world,
                           "api": "rmt_app_b200.rmtExeBatchN2Sharded(modelInput, sweep, zNo=200, tNo=5): every rank integrates its "
                                  "contiguous block; device time (CUDA events, max over ranks) incl. H2D of the rank's inputs, setup, "
                                  "integrator, ONE packed all-gather (status + tails; with gather='final' also the [n][200] profiles)",
                           "reference_rhs_200_nodes_s_per_call": (cpu_file or {}).get("config5_modelEquationN2_200_nodes_s_per_call")}}
        if e2:
            n2["ensemble"]["ncu"] = dict(e2, calibration_stale=stale2)
        del psw2, sw2, r2g
        # the stage-pipelined kernel on ensembles that fill its rounds (148 blocks x 64 reactors per GPU): 200 nodes, one
        # round; 50 nodes, six rounds — what engine.compile_model_n2 picks for these sizes
        pipe = {}
        for tag, Bp_, zp_ in (("one_round_200_nodes", 9472, 200), ("50000_x_50_nodes", 50000, 50)):
            cmp2 = engine.compile_model_n2(mi2, Bp_, zp_)
            swp = cases.config3_sweep(Bp_, 20240613 + rank)
            dswp = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in swp.items()}
            wsp2 = engine.Workspace()
            engine.n2_solve_ensemble(cmp2, mi2, dswp, Bp_, zNo=zp_, tNo=5, period=0.5, keep_on_device=True, workspace=wsp2)
            barrier()
            p0 = torch.cuda.Event(enable_timing=True); p1 = torch.cuda.Event(enable_timing=True)
            p0.record()
            rp2 = engine.n2_solve_ensemble(cmp2, mi2, dswp, Bp_, zNo=zp_, tNo=5, period=0.5, keep_on_device=True, workspace=wsp2)
            p1.record(); torch.cuda.synchronize()
            sp = max_over_ranks(p0.elapsed_time(p1)*1e-3)
            attp, okp = sum_over_ranks([float(rp2.stats[3].double().sum().item()), float((rp2.status == 0).sum().item())])
            ip = cmp2.load(local_rank).info
            # per attempt and node: f + Jacobian blocks, (s-1) RHS, one LU, s LU solves and the stage combinations
            linp = (2.0*nn**3)/3.0 + ip.stages*(2.0*nn*nn + 4.0*nn*ip.stages)
            algp = attp*zp_*(ip.flops_jac_alg + (ip.stages - 1)*ip.flops_rhs_alg + linp)
            ep, stalep = calibration_for(cal, "rmt_n2_solve_pipeline", cmp2.key())
            pipe[tag] = {"instances_per_gpu": Bp_, "nodes": zp_, "seconds": sp, "instances_per_s": world*Bp_/sp, "converged": int(okp),
                         "kernel": "stage pipeline" if cmp2.lanes == 0 else "lanes (%d per reactor)" % cmp2.lanes, "block": cmp2.block,
                         "node_rhs_evals_per_s": attp*zp_*ip.stages/sp, "fp64_tflops_algorithmic": algp/sp/1e12,
                         "fp64_frac_of_measured_peak": algp/sp/1e12/fp64_peak/world}
            if ep and zp_ == 200:
                pipe[tag]["ncu"] = dict(ep, calibration_stale=stalep)
            del dswp, wsp2, rp2
        n2["stage_pipeline"] = pipe

    if rank == 0:
        steps = args.steps
        total = world*B*steps
        solve_s = solve_ms*1e-3
        e1, stale1 = calibration_for(cal, "rmt_n1_solve", cm.key())
        er, staler = calibration_for(cal, "rmt_n1_rhs", cm.key())
        ej, stalej = calibration_for(cal, "rmt_n1_jac", cm.key())
        roof = {
            "kernel": "rmt_n1_solve", "bound": "fp64", "achieved": alg/solve_s/1e12, "peak": fp64_peak,
            "unit": "TFLOP/s", "frac": alg/solve_s/1e12/fp64_peak,
            "traffic": None, "traffic_algorithmic": 8.0*(22 + n)*B + 20.0*B,
            "achieved_weighted": wt/solve_s/1e12, "frac_weighted": wt/solve_s/1e12/fp64_peak,
            "peak_source": "rmt_dfma_peak measured in this run (MEASURED_PEAKS.json has no FP64 figure; "
                           "nominal 148 SM x 64 FMA/clk x 2 x 1.965 GHz = 37.2)",
            "kernel_ms": solve_ms, "kernel_share_of_step": solve_ms/(elapsed_ms/steps),
            "flops_per_launch_alg": alg, "flops_per_launch_weighted": wt,
            "attempts_per_solve": att/B, "rhs_evals_per_solve": (nfev + att)/B,
            "note": "contract offers hbm|tensor; this kernel keeps all state on-chip (HBM traffic = 200 B in + 64 B "
                    "out per reactor) and has no contraction, so the bounding pipe is FP64 FMA",
        }
        if e1:
            # numbers read off the ncu capture named in the calibration file (per launch of 2^20 reactors / per step attempt)
            fl = 2*e1["dfma_per_attempt"] + e1["dmul_per_attempt"] + e1["dadd_per_attempt"]
            roof.update(traffic=e1["dram_bytes_per_reactor"]*B, achieved_executed=fl*att/solve_s/1e12,
                        frac_executed=fl*att/solve_s/1e12/fp64_peak, fp64_pipe_busy_ncu=e1["fp64_pipe_busy"],
                        calibration={"file": "profiles/r02_calibration.json", "source": e1.get("source"),
                                     "module_key": e1.get("module_key")}, calibration_stale=stale1)
        line = {
            "metric": METRIC, "value": total/(elapsed_ms*1e-3), "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms/steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(world),
            "clocks": clocks,
            "e2e": {"value": world*B*steps/e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "api": "rmt_app_b200.rmtExeBatch(modelInput, sweep, workspace=...) with pinned host tensors in, pinned host arrays out",
                    "converged": e2e_ok,
                    "c_abi_host_call": {"value": world*B*steps/cabi_s, "unit": UNIT, "converged": cabi_ok,
                                        "api": "rmt_n1_solve_host (include/rmt_b200.h) with pinned host buffers, same bytes"}},
            "gpu_launches": 2*steps,
            "converged": n_ok_all, "instances": world*B,
            "roofline": roof,
            "roofline_rhs_kernel": {
                "kernel": "rmt_n1_rhs", "bound": "hbm", "achieved": rhs_bytes/(rhs_ms*1e-3)/1e9, "peak": hbm_peak,
                "unit": "GB/s", "frac": rhs_bytes/(rhs_ms*1e-3)/1e9/hbm_peak,
                "traffic": er["dram_bytes_per_launch"]*B/er["instances"] if er else None,
                "frac_by_traffic": (er["dram_bytes_per_launch"]*B/er["instances"]/(rhs_ms*1e-3)/1e9/hbm_peak) if er else None,
                "fp64_pipe_busy_ncu": er["fp64_pipe_busy"] if er else None, "calibration_stale": staler,
                "peak_source": hbm_src, "kernel_ms": rhs_ms, "rhs_evals_per_s": B/(rhs_ms*1e-3),
                "timing": "CUDA events, L2 flushed (256 MB memset) before every launch",
                "fp64_tflops_weighted": B*info.flops_rhs_wt/(rhs_ms*1e-3)/1e12,
            },
            "roofline_jac_kernel": {
                "kernel": "rmt_n1_jac", "bound": "hbm", "achieved": jac_bytes/(jac_ms*1e-3)/1e9, "peak": hbm_peak,
                "unit": "GB/s", "frac": jac_bytes/(jac_ms*1e-3)/1e9/hbm_peak,
                "traffic": ej["dram_bytes_per_launch"]*B/ej["instances"] if ej else None,
                "frac_by_traffic": (ej["dram_bytes_per_launch"]*B/ej["instances"]/(jac_ms*1e-3)/1e9/hbm_peak) if ej else None,
                "fp64_pipe_busy_ncu": ej["fp64_pipe_busy"] if ej else None, "calibration_stale": stalej,
                "peak_source": hbm_src, "kernel_ms": jac_ms, "jac_evals_per_s": B/(jac_ms*1e-3),
                "timing": "CUDA events, L2 flushed (256 MB memset) before every launch",
                "fp64_tflops_weighted": B*info.flops_jac_wt/(jac_ms*1e-3)/1e12,
            },
            "rhs_evals_per_sec_in_solver": world*(nfev + att)/solve_s,
            "setup_kernel_ms": setup_ms,
        }
        if cpu_single:
            line["reference_rhs_only"] = {"modelEquationN1_evals_per_s_per_core": cpu_single["modelEquationN1_evals_per_s_per_core"],
                                          "s_per_call": cpu_single["modelEquationN1_s_per_call"],
                                          "what": "the unmodified reference's RHS called on states of its own solution, one host core"}
        for key, val in (("tight_tolerance", tight), ("config1_single_reactor", config1), ("n2_dynamic_model", n2),
                         ("config4_population", config4), ("strong_scaling", strong), ("profile_101_points", profile),
                         ("cpu_baseline", cpu_baseline)):
            if val is not None:
                line[key] = val
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _timed(fn):
    t0 = time.perf_counter()
    fn()
    return time.perf_counter() - t0


_JSON_OUT = None


def protect_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on
    stdout at any NCCL_DEBUG level, NVRTC / the driver may warn), so keep a private duplicate of the real stdout
    for the JSON line and point file descriptor 1 at stderr for everything else."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--instances", type=int, default=B_PER_GPU, help="reactors per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-baselines", action="store_true",
                    help="only the long CPU measurements of BASELINE.md section 3 (reference on the host cores; minutes)")
    ap.add_argument("--no-n2", action="store_true", help="skip the informational N2 (dynamic model) timings")
    ap.add_argument("--no-config4", action="store_true", help="skip the informational parameter-estimation population timing")
    ap.add_argument("--no-profile", action="store_true", help="skip the informational 101-point-profile timing")
    ap.add_argument("--no-tight", action="store_true", help="skip the parity-grade-tolerance timing (Rodas4, rtol 1e-9)")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling block (one fixed ensemble over all ranks)")
    ap.add_argument("--strong-instances", type=int, default=1 << 23)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1 and args.impl != "reference":
        # convenience: relaunch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    protect_stdout()
    if args.cpu_baselines:
        if rank == 0:
            run_cpu_baselines(args)
        return
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
    else:
        run_gpu_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
