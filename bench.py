#!/usr/bin/env python
"""Benchmark of the TriTD-ADMM hot path (BASELINE.json metric: ADMM iterations/s, fp64).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfgX] [--impl reference]

One "step" = one ADMM iteration (triple_decomp_ADMM.m:31-66) on the synthetic tensor of the named BASELINE config.
N=1 default workload: cfg3 (240x320x300, r=5), the shape the metric is quoted on.  N>1 (launched by
torch.distributed.run, one rank per GPU): the SAME tensor sharded along mode 3 over the ranks ("strong" scaling); the
[RHS ; Gram] partials travel through NVLink peer mailboxes inside the update kernels (NCCL only for set-up).

Prints ONE JSON line (rank 0):
  value     iterations/s with all inputs resident in HBM, timed with CUDA events on the launching stream over
            exactly K iterations after W warm-up iterations, max over ranks; at N>1 an untimed device-side
            rendezvous (1-element NCCL all-reduce on the timed stream) precedes the first event, so process start
            skew is not billed to the K steps;
  e2e       iterations/s of the reference-facing call (tritd_admm_f64 through ctypes) with HOST buffers: pinned
            D in, A/B/C/O/errHist out, H2D and D2H inside the timed region;
  roofline  k_admm, the fused element-wise kernel: 48*N algorithmic bytes per launch (3 reads D, Y_L, Z + 3 writes
            T', Y_L, Z; the sparse pair (E, Y_O) is kept as the one array Z = R3) / its CUDA-event time / the measured
            HBM peak; `survey_yardstick_64N` restates the same time against SURVEY 8d's 4-read/4-write figure; `ppass`
            and `fp64` carry the FP64-tensor fractions against the DMMA peak measured in this run;
  parity_vs_fixture  max relative deviation of the first errHist values of this very run (any N) from the committed
            CPU-oracle fixture tests/golden/fullsize_errhist.json;
  time_to_tol  the reference's own call (tol 1e-5, maxIter 100) timed from entry until the outputs are on the host;
  final_rre (N=1) outcome of that full run on the headline config and on cfg1: RRE = ||triple_product(A,B,C) - L0|| / ||L0||
            (the drivers' evaluate(), traffic_triple_comparison.m:194-199; formed on the device), iteration count and last
            errHist value, next to the CPU oracle's numbers for the same run (fixture tests/golden/final_rre.json);
  cpu_baseline  the multi-threaded CPU port of the oracle timed on this box's host cores (rank 0, N=1);
  configs   (N=1) device-resident iterations/s of the other BASELINE configs; secondary = cfg5, the shape the north
            star's scaling target names (every N).
--impl reference times the CPU port alone (MATLAB cannot run here; with octave/matlab on PATH and
TRITD_REFERENCE_DIR set, the reference's own .m is timed through tools/reference_dump.m instead).
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, "triple-tensor-decomposition-with-admm_b200"),):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

# k_admm moves 3 reads (D, Y_L, Z) + 3 writes (T', Y_L, Z) = 48 bytes per element: O is not stored inside the loop
# and the sparse pair (E, Y_O) travels as the one array Z = R3.  SURVEY 8d's figure for the straightforward
# fused kernel is 64*N (4 reads D, Y_L, E, Y_O + 4 writes); it is reported next to it as the yardstick.
FUSED_BYTES = 48.0
SURVEY_FUSED_BYTES = 64.0
METRIC = "admm_iterations_per_second"
UNIT = "iter/s"
L2_MB = 126.0


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_fixture():
    try:
        with open(os.path.join(ROOT, "tests", "golden", "fullsize_errhist.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def make_config_dict(name, world):
    """`config` of the JSON line -- built by the same function in both arms so the driver sees identical dicts."""
    from tritd import synth
    n1, n2, n3, r = synth.CONFIGS[name][:4]
    n3l = -(-n3 // world)
    mb = n1 * n2 * n3l * 8e-6
    if 4 * mb > 2 * L2_MB:
        l2 = "inputs larger than L2 (4 streamed state arrays (D, Y_L, Z, T) x %.0f MB per rank vs %.0f MB L2), no flush" % (mb, L2_MB)
    else:
        l2 = ("per-rank state (4 streamed arrays x %.0f MB) is comparable to the %.0f MB L2: iterations run back to back exactly as "
              "in a real solve, the state an iteration leaves in L2 is what the next one finds; no flush" % (mb, L2_MB))
    return {"workload": f"{name}: {synth.DESCRIPTIONS[name]}", "n1": n1, "n2": n2, "n3": n3, "r": r,
            "sharding": f"mode-3 slabs over {world} rank(s)", "l2": l2}


def workload_arrays(name, t0=None, t1=None):
    """Host arrays of the named config; [t0,t1) selects a mode-3 slab."""
    from tritd import synth
    n1, n2, n3, r, kind, frac, seed, opts = synth.CONFIGS[name]
    if kind == "lowrank_sparse":
        D = synth.make_lowrank_sparse(n1, n2, n3, r, frac, seed, t0=t0 or 0, t1=t1)
    else:
        D = synth.make_config(name)["D"]
        if t0 is not None:
            D = np.asfortranarray(D[:, :, t0:t1])
    A0, B0, C0 = synth.init_factors(n1, n2, n3, r, 100 + seed)
    if t0 is not None:
        C0 = np.asfortranarray(C0[:, :, t0:t1])
    return D, r, dict(opts), A0, B0, C0, (n1, n2, n3)


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    NAMES = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
             0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
             0x100: "display_clock_setting"}

    def run(self):
        if not self.ok:
            return
        while not self.stop_flag:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in self.NAMES.items():
                    if r & bit and nm != "gpu_idle":
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.002)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def cpu_oracle_rate(name, steps, warmup, budget_s):
    """ADMM iterations/s of the CPU port on the host cores: `warmup` untimed + `steps` timed iterations (time stamps
    taken inside the loop).  The port is oracle/tritd_oracle_mt.py -- the oracle's statements and passes with every
    N-sized operation on all host threads, like MATLAB's multi-threaded built-ins (checked against the numpy oracle in
    tests/test_oracle.py).  To bound the run the tensor is cut to its first `t_slices` mode-3 slices -- every pass of
    the iteration is linear in n3 -- and the rate is scaled by t_slices/n3 to the full workload."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import tritd_oracle_mt as mt
    from tritd import synth
    cores = os.cpu_count() or 1
    n1, n2, n3, r, kind, frac, seed, opts = synth.CONFIGS[name]
    cal = min(4, n3)
    big = n1 * n2 * n3 > 2e8           # cfg5: never materialise the whole tensor on the host for a CPU sample
    Dfull = None
    if big:
        first = lambda k: synth.make_lowrank_sparse(n1, n2, n3, r, frac, seed, t0=0, t1=k)   # noqa: E731
    else:
        Dfull = workload_arrays(name)[0]
        first = lambda k: np.asfortranarray(Dfull[:, :, :k])   # noqa: E731
    A0, B0, C0 = synth.init_factors(n1, n2, n3, r, 100 + seed)
    t = time.perf_counter()
    mt.triple_decomp_ADMM(first(cal), r, dict(opts, maxIter=2, tol=0.0, disp=0), A0, B0, np.asfortranarray(C0[:, :, :cal]), threads=cores)
    per_slice_iter = (time.perf_counter() - t) / (2 * cal)
    t_slices = int(max(min(4, n3), min(n3, budget_s / max(per_slice_iter * (steps + warmup), 1e-9))))
    Ds = first(t_slices); Cs = np.asfortranarray(C0[:, :, :t_slices])
    stamps = [time.perf_counter()]
    mt.triple_decomp_ADMM(Ds, r, dict(opts, maxIter=steps + warmup, tol=0.0, disp=0), A0, B0, Cs,
                          on_iter=lambda *a: stamps.append(time.perf_counter()), threads=cores)
    dt = stamps[-1] - stamps[warmup]
    rate_full = steps / dt * (t_slices / n3)
    sample = (f"{warmup}+{steps} iterations of the multi-threaded CPU port (oracle/tritd_oracle_mt.py: the oracle's statements on "
              f"torch CPU tensors, {cores} threads for every N-sized pass and dgemm) on the first {t_slices} of {n3} mode-3 slices of "
              f"{name} ({n1}x{n2}x{t_slices}); rate scaled by {t_slices}/{n3} (every pass is linear in n3)")
    return rate_full, sample, dt


def matlab_reference_rate(name, steps):
    """The reference's own .m under Octave / MATLAB, when an interpreter is on PATH and TRITD_REFERENCE_DIR points at
    a checkout of the reference (it does not exist on the GPU box by default): tools/reference_dump.m with
    maxIter = steps.  Returns (rate, sample, interpreter) or None."""
    exe = shutil.which("octave") or shutil.which("matlab")
    refdir = os.environ.get("TRITD_REFERENCE_DIR")
    if not exe or not refdir or not os.path.isdir(os.path.join(refdir, "fast_robust_triple_tensor")):
        return None
    from scipy.io import loadmat, savemat
    D, r, opts, A0, B0, C0, shape = workload_arrays(name)
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.mat"), os.path.join(td, "out.mat")
        o = {k: float(opts[k]) for k in ("mu", "rho", "lambda", "lambda2")}
        o.update(maxIter=float(steps), tol=0.0, disp=0.0)
        savemat(fin, dict(D=D, r=float(r), opts=o, A0=A0, B0=B0, C0=C0), format="5")
        call = f"addpath('{os.path.join(ROOT, 'tools')}'); reference_dump('{refdir}', '{fin}', '{fout}');"
        cmd = [exe, "--eval", call] if exe.endswith("octave") else [exe, "-batch", call]
        subprocess.run(cmd, check=True, timeout=1800, stdout=subprocess.DEVNULL)
        m = loadmat(fout)
    sec = float(m["seconds"].ravel()[0])
    its = int(m["errHist"].size)
    return its / sec, f"{its} iterations of the unmodified fast_robust_triple_tensor/triple_decomp_ADMM.m under {os.path.basename(exe)} on {name}", exe


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload
    steps, warm = args.steps, args.warmup
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    kind, note = "port", "reference is MATLAB and cannot run here (no MATLAB/Octave); this is the multi-threaded CPU port of the oracle"
    real = None
    try:
        real = matlab_reference_rate(name, max(2, min(steps, 5)))
    except Exception as exc:
        note += f" (an Octave/MATLAB attempt failed: {type(exc).__name__})"
    if real:
        rate, sample, exe = real
        kind, note = "reference", f"the reference's own .m timed under {exe}"
    else:
        rate, sample, dt = cpu_oracle_rate(name, steps, warm, budget_s=120.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": 1e3 / rate, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": make_config_dict(name, max(world, args.gpus, 1)),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": os.cpu_count(), "kind": kind, "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": note,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="cfg3")
    ap.add_argument("--impl", default="tritd", choices=["tritd", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--secondary", default="cfg5", help="second workload measured device-resident only (\"\" to skip)")
    ap.add_argument("--configs", default="cfg1,cfg2,cfg4", help="other BASELINE configs measured device-resident at N=1 (\"\" to skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import tritd
    from tritd import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (tritd has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    name = args.workload
    n1, n2, n3, r = synth.CONFIGS[name][:4]
    t0, t1 = tritd.slab_bounds(n3, world, rank) if world > 1 else (0, n3)
    D, r, opts, A0, B0, C0, _ = workload_arrays(name, t0 if world > 1 else None, t1 if world > 1 else None)
    N_global = n1 * n2 * n3
    K, W = args.steps, args.warmup
    fixture = load_fixture()

    if world > 1:
        idbuf = [tritd.Context.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(idbuf, src=0)
        ctx = tritd.Context(local, rank, world, idbuf[0])
    else:
        ctx = tritd.Context(local)
    # the library launches on THIS stream, and so do the timing events below (a dedicated non-default
    # stream: the C ABI treats a NULL stream handle as "use the context's own stream")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)
    rdv = torch.zeros(1, dtype=torch.float64, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def device_rendezvous():
        """N>1: an untimed 1-element all-reduce on the timed stream -- every rank's stream waits here for the slowest
        process, so the first timed exchange does not pay the processes' start skew."""
        if world > 1:
            dist.all_reduce(rdv)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    dmma_peak = max_over_ranks(-ctx.measure_dmma_peak(60.0))
    dmma_peak = -dmma_peak              # min over ranks (conservative denominator)

    def parity(wname, eh):
        f = fixture.get(wname)
        if not f:
            return None
        k = min(len(f["errHist"]), len(eh))
        ref = np.array(f["errHist"][:k])
        return {"max_rel_dev": float(np.max(np.abs(np.asarray(eh[:k]) - ref) / ref)), "iterations": k, "tolerance": 1e-8,
                "fixture": "tests/golden/fullsize_errhist.json (CPU oracle)"}

    def measure(wname, Dw, A0w, B0w, C0w, optsw, shape, rw, Kw, Ww, sample_clocks):
        """Device-resident throughput of `Kw` iterations after `Ww` warm-up iterations (+ per-kernel times)."""
        m1, m2, m3 = shape
        t0w, t1w = tritd.slab_bounds(m3, world, rank) if world > 1 else (0, m3)
        m3l = t1w - t0w
        bo = dict(optsw, maxIter=Ww + Kw + 8, tol=0.0, disp=0)     # tol=0: the stopping rule never fires, every step is real
        prob = tritd.Problem(ctx, m1, m2, m3l, rw)
        prob.set_D(Dw)
        prob.init(bo, A0w, B0w, C0w)
        prob.enqueue(Ww)
        prob.sync()
        barrier()
        sampler = ClockSampler(local) if sample_clocks else None
        if sampler:
            sampler.start()
        launches0 = ctx.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        device_rendezvous()
        e0.record(stream)
        prob.enqueue(Kw)                      # Kw iterations back to back on `stream` (CUDA-graph replays), no host sync
        e1.record(stream)
        barrier()
        if sampler:
            sampler.stop_flag = True
            sampler.join()
        ms_total = max_over_ranks(e0.elapsed_time(e1))
        launches = ctx.launches - launches0
        prob.sync()
        res = prob.get(want_O=False)
        assert res["iters"] == Ww + Kw, (res["iters"], Ww + Kw)
        assert np.all(np.isfinite(res["errHist"]))
        # second pass over the same steps with the library's per-phase CUDA events (plain launches) for the
        # per-kernel times the roofline needs; kept out of `value` because the events add gaps
        prob.init(bo, A0w, B0w, C0w)
        prob.enqueue(Ww)
        prob.sync()
        barrier()
        prob.set_profiling(True)
        device_rendezvous()
        prob.enqueue(min(Kw, 50))
        phase_ms, nprof = prob.phase_ms()
        prob.set_profiling(False)
        prob.close()
        Nl = m1 * m2 * m3l
        RSw = (rw * rw + 7) // 8 * 8
        fused_ms = phase_ms[4] / max(1, nprof)
        ppass_ms = phase_ms[2] / max(1, nprof)
        return dict(value=Kw / (ms_total * 1e-3), ms_total=ms_total, launches=int(launches), N_local=Nl, N_global=m1 * m2 * m3,
                    fused_ms=fused_ms, achieved=FUSED_BYTES * Nl / (fused_ms * 1e-3) * 1e-9,
                    ppass_ms=ppass_ms, ppass_tflops=2.0 * Nl * RSw / (ppass_ms * 1e-3) * 1e-12,
                    fused_tflops=(2.0 * Nl * 4 * ((rw * rw + 3) // 4) + 2.0 * Nl * RSw) / (fused_ms * 1e-3) * 1e-12,
                    phase_ms={nm: phase_ms[i] / max(1, nprof) for i, nm in enumerate(tritd.PHASES)},
                    clocks=sampler.summary() if sampler else None, state_mb=Nl * 8e-6, parity=parity(wname, res["errHist"]))

    # ---------------- device-resident throughput ----------------
    m = measure(name, D, A0, B0, C0, opts, (n1, n2, n3), r, K, W, True)
    value, ms_total, launches, N_local = m["value"], m["ms_total"], m["launches"], m["N_local"]

    # ---------------- roofline of the dominant kernel (fused element-wise pass) ----------------
    peak, peak_src = load_peaks()
    fused_ms, achieved = m["fused_ms"], m["achieved"]
    R = r * r
    flops_iter = 8.0 * N_global * R + 2.0 * R * R * (n2 * n3 + n1 * n3 + n1 * n2)
    roofline = {"bound": "hbm", "kernel": "k_admm (TMA in / DMMA L reconstruction + O/E/dual/T update on the state D, Y_L, Z + residual norms + next mode-1 MTTKRP / TMA out)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "traffic_source": "not measured in this run; one ncu --set full capture per change is committed under profiles/ "
                                  "(dram__bytes_read.sum + dram__bytes_write.sum per launch: profiles/r02_final_cfg3_ncu_summary.csv, "
                                  "590.5 MB + 502.0 MB for the cfg3 launch of this kernel; r02_final_cfg5slab_ncu_summary.csv)",
                "peak_source": peak_src, "algorithmic_bytes_per_launch": FUSED_BYTES * N_local, "kernel_ms": fused_ms,
                "kernel_ms_source": "CUDA events recorded by the library around k_admm on the launching stream, averaged over a second pass of the same steps",
                "survey_yardstick_64N": {"GBps": SURVEY_FUSED_BYTES * N_local / (fused_ms * 1e-3) * 1e-9,
                                         "frac_of_peak": SURVEY_FUSED_BYTES * N_local / (fused_ms * 1e-3) * 1e-9 / peak,
                                         "note": "the same kernel time against SURVEY 8d's 64*N (a kernel that keeps E and Y_O as two arrays "
                                                 "must move that much); > 1 means faster than that formulation's HBM roofline"},
                "iteration_GBps_vs_96N": 96.0 * N_global / world / (ms_total / K * 1e-3) * 1e-9,
                "iteration_GBps_real_56N": 56.0 * N_global / world / (ms_total / K * 1e-3) * 1e-9,
                "fp64": {"dmma_peak_TFLOPs": dmma_peak, "dmma_peak_source": "measured in this run (tritd_measure_dmma_peak, 60 ms of DMMA.8x8x4 probes; "
                                                                            "DMMA and DFMA share one FP64 datapath on B200, profiles/r02_microbench.log)",
                         "iteration_TFLOPs": flops_iter / world / (ms_total / K * 1e-3) * 1e-12,
                         "iteration_frac_of_dmma_peak": flops_iter / world / (ms_total / K * 1e-3) * 1e-12 / dmma_peak,
                         "k_admm_dmma_TFLOPs": m["fused_tflops"], "k_admm_frac_of_dmma_peak": m["fused_tflops"] / dmma_peak},
                "ppass": {"bound": "tensor", "kernel": "k_ppass (P = T x_1 A1: FP64 DMMA fed by TMA, shared by update_B and update_C)",
                          "achieved": m["ppass_tflops"], "peak": dmma_peak, "unit": "TFLOP/s", "frac": m["ppass_tflops"] / dmma_peak,
                          "algorithmic_flops_per_launch": 2.0 * N_local * ((R + 7) // 8 * 8), "kernel_ms": m["ppass_ms"]},
                "phase_ms_per_iter": m["phase_ms"]}

    # ---------------- end to end through the reference-facing call, host buffers ----------------
    e2e = None
    ttt = None
    if not args.no_e2e:
        Dp = torch.empty(D.size, dtype=torch.float64).pin_memory()
        Dn = Dp.numpy().reshape(D.shape, order="F")
        Dn[...] = D
        Op = torch.empty(D.size, dtype=torch.float64).pin_memory()         # the caller's O buffer, pinned as well
        On = Op.numpy().reshape(D.shape, order="F")
        # "the call a user makes": one solver call runs opts.maxIter (= 100, the reference's setting) iterations and
        # moves D in and A, B, C, O, errHist out; K steps = ceil(K / maxIter) such calls back to back, every one with
        # its own H2D and D2H inside the timed region (tol = 0 so each call runs all its iterations)
        per_call = int(opts["maxIter"])
        ncalls = max(1, -(-K // per_call))
        e2e_opts = dict(opts, maxIter=per_call, tol=0.0, disp=0)
        tritd.triple_decomp_ADMM(Dn, r, dict(e2e_opts, maxIter=3), A0, B0, C0, ctx=ctx)       # warm-up call
        barrier()
        tw = time.perf_counter()
        infos = []
        for _ in range(ncalls):
            A, B, C, O, eh, info = tritd.triple_decomp_ADMM(Dn, r, e2e_opts, A0, B0, C0, ctx=ctx, return_info=True, out_O=On)
            assert len(eh) == per_call
            infos.append(info)
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - tw)
        h2d = (D.nbytes + A0.nbytes + B0.nbytes + C0.nbytes) / per_call
        d2h = (O.nbytes + A.nbytes + B.nbytes + C.nbytes + eh.nbytes) / per_call
        e2e = {"value": ncalls * per_call / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "call": f"{ncalls} x tritd_admm_f64 of {per_call} iterations each (the reference's maxIter; pinned host D and O; "
                       "H2D of D/A0/B0/C0 + iterations + D2H of A,B,C,O,errHist inside every call)",
               "iterations": ncalls * per_call, "seconds": dt,
               "h2d_ms_per_call": float(np.mean([i["h2d_ms"] for i in infos])),
               "iterate_ms_per_call": float(np.mean([i["iterate_ms"] for i in infos])),
               "d2h_ms_per_call": float(np.mean([i["d2h_ms"] for i in infos])),
               "parity_vs_fixture": parity(name, eh)}
        # time-to-tolerance with the reference's own options (tol 1e-5, maxIter 100)
        barrier()
        tw = time.perf_counter()
        out = tritd.triple_decomp_ADMM(Dn, r, dict(opts, disp=0), A0, B0, C0, ctx=ctx, return_info=True, out_O=On)
        torch.cuda.synchronize()
        dt2 = max_over_ranks(time.perf_counter() - tw)
        nit = len(out[4])
        ttt = {"seconds_host_buffers": dt2, "seconds_device_loop": out[5]["iterate_ms"] * 1e-3, "iterations": nit,
               "tol": opts["tol"], "maxIter": opts["maxIter"], "final_errHist": float(out[4][-1]),
               "stopped_by": "tol" if nit < int(opts["maxIter"]) else "maxIter (the relative-change rule did not fire at the reference's tol: this is time to maxIter)"}
        ctx.trim()
        del Dp, Op, Dn, On
        if world == 1:
            # a case where the rule does fire (the golden 'stop' case: cfg1-like 30^3, r = 3, tol 2e-2)
            ws = synth.make_config("cfg1", shrink=(30, 30, 30))
            so = dict(ws["opts"], maxIter=100, tol=2e-2, disp=0)
            f0 = synth.init_factors(30, 30, 30, 3, 101)
            tritd.triple_decomp_ADMM(ws["D"], 3, so, *f0, ctx=ctx)
            tw = time.perf_counter()
            o2 = tritd.triple_decomp_ADMM(ws["D"], 3, so, *f0, ctx=ctx)
            ttt["case_where_the_rule_fires"] = {"workload": "cfg1-like 30x30x30, r=3, tol=2e-2", "iterations": len(o2[4]),
                                                "seconds_host_buffers": time.perf_counter() - tw}
            ctx.trim()

    # ---------------- the other BASELINE configs, device-resident (N=1) ----------------
    others = []
    del D
    if world == 1 and args.configs:
        for cname in [c for c in args.configs.split(",") if c and c != name and c != args.secondary]:
            try:
                Dc, rc, oc, a0, b0, c0, shp = workload_arrays(cname)
                Kc = max(10, min(K, 200 if np.prod(shp) < 5e7 else 40))
                mc = measure(cname, Dc, a0, b0, c0, oc, shp, rc, Kc, 5, False)
                others.append({"workload": f"{cname}: {synth.DESCRIPTIONS[cname]}", "value": mc["value"], "unit": UNIT, "steps": Kc,
                               "ms_per_step": mc["ms_total"] / Kc, "k_admm_GBps": mc["achieved"], "roofline_frac_k_admm": mc["achieved"] / peak,
                               "ppass_frac_of_dmma_peak": mc["ppass_tflops"] / dmma_peak, "phase_ms_per_iter": mc["phase_ms"],
                               "parity_vs_fixture": mc["parity"]})
                del Dc
            except Exception as exc:
                others.append({"workload": cname, "error": f"{type(exc).__name__}: {exc}"})

    # ---------------- secondary workload: cfg5, the shape the north star's scaling target names ----------------
    secondary = None
    if args.secondary and args.secondary != name:
        sname = args.secondary
        s1, s2, s3, sr = synth.CONFIGS[sname][:4]
        st0, st1 = tritd.slab_bounds(s3, world, rank) if world > 1 else (0, s3)
        try:
            Ds, sr, sopts, sA0, sB0, sC0, _ = workload_arrays(sname, st0 if world > 1 else None, st1 if world > 1 else None)
            Ks = max(5, min(K, 30))
            ms_ = measure(sname, Ds, sA0, sB0, sC0, sopts, (s1, s2, s3), sr, Ks, 3, False)
            secondary = {"workload": f"{sname}: {synth.DESCRIPTIONS[sname]}", "metric": METRIC, "value": ms_["value"], "unit": UNIT,
                         "steps": Ks, "warmup": 3, "ms_per_step": ms_["ms_total"] / Ks, "scaling": "strong",
                         "roofline_frac_k_admm": ms_["achieved"] / peak, "k_admm_GBps": ms_["achieved"],
                         "k_admm_frac_of_dmma_peak": ms_["fused_tflops"] / dmma_peak, "ppass_frac_of_dmma_peak": ms_["ppass_tflops"] / dmma_peak,
                         "hbm_bound_iters_per_s_96N": peak * 1e9 / (96.0 * ms_["N_global"] / world),
                         "phase_ms_per_iter": ms_["phase_ms"], "state_MB_per_array_per_rank": ms_["state_mb"]}
        except Exception as exc:           # the headline line must not depend on the secondary workload
            if world > 1:
                raise                      # (a rank dropping out of a sharded solve would stall the others)
            secondary = {"workload": sname, "error": f"{type(exc).__name__}: {exc}"}
        Ds = None
        del Ds

    # ---------------- CPU baseline beside it (rank 0, N=1 only) ----------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, sample, _ = cpu_oracle_rate(name, steps=3, warmup=1, budget_s=20.0)
        cpu = {"value": rate, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": sample}
        if secondary and "error" not in secondary:
            try:
                rate5, sample5, _ = cpu_oracle_rate(args.secondary, steps=2, warmup=1, budget_s=12.0)
                secondary["cpu_baseline"] = {"value": rate5, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": sample5}
            except Exception as exc:
                secondary["cpu_baseline"] = {"skipped": f"{type(exc).__name__}: {exc}"}

    # ---------------- final RRE of the reference's full run (north_star: "the final RRE reported"), N=1 ----------------
    # Last thing before the line is printed, every case on its own: nothing above can be affected by a failure here.
    # The CPU oracle's numbers for the same runs come from a committed fixture (tests/golden/make_final_rre.py).
    final_rre = None
    if world == 1 and not args.no_e2e:
        final_rre = []
        try:
            with open(os.path.join(ROOT, "tests", "golden", "final_rre.json")) as f:
                rre_fixture = json.load(f)
        except Exception:
            rre_fixture = {}
        for cname in dict.fromkeys((name, "cfg1")):
            try:
                wt = synth.make_config(cname, with_truth=True)          # the seeded data of the timed runs + the low-rank part it was built from
                fo = dict(wt["opts"], disp=0)                            # the reference's own options for this kind of data
                fA, fB, fC, fO, feh = tritd.triple_decomp_ADMM(wt["D"], wt["r"], fo, wt["A0"], wt["B0"], wt["C0"], ctx=ctx)
                rmse, nrmse = tritd.evaluate(fA, fB, fC, wt["L0"], ctx=ctx)
                ent = {"workload": f"{cname}: {synth.DESCRIPTIONS[cname]}", "RRE": nrmse, "rmse": rmse, "iterations": len(feh),
                       "final_errHist": float(feh[-1]), "tol": fo["tol"], "maxIter": int(fo["maxIter"]),
                       "definition": "nrmse of evaluate(triple_product(A,B,C), L0, all-true mask) = ||Xhat - L0||_F / ||L0||_F "
                                     "(traffic_triple_comparison.m:194-199) after [A,B,C,O,errHist] = triple_decomp_ADMM(D, r, opts) with the "
                                     "reference's options; L0 = the low-rank part the synthetic D was built from; reconstruction and norms on the device"}
                fx = rre_fixture.get(cname)
                if fx and fx.get("shape") == list(wt["shape"]):
                    ent["cpu_oracle"] = {"RRE": fx["RRE"], "iterations": fx["iterations"], "final_errHist": fx["final_errHist"],
                                         "source": "tests/golden/final_rre.json (CPU oracle, same inputs and options)"}
                    ent["same_iteration_count_as_cpu_oracle"] = bool(fx["iterations"] == len(feh))
                    ent["RRE_abs_dev_from_cpu_oracle"] = abs(nrmse - fx["RRE"])
                final_rre.append(ent)
                ctx.trim()
            except Exception as exc:
                final_rre.append({"workload": cname, "error": f"{type(exc).__name__}: {exc}"})

    if rank == 0:
        cfg = make_config_dict(name, world)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg,
            "opts": {k: opts[k] for k in ("mu", "rho", "lambda", "lambda2")},
            "clocks": m["clocks"], "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "parity_vs_fixture": m["parity"], "time_to_tol": ttt, "final_rre": final_rre, "configs": others, "secondary": secondary,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
