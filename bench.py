#!/usr/bin/env python
"""bench.py -- headline benchmark of the filtering hot path (BASELINE.json metric) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (config.workload): BASELINE.json configs[2] -- bootstrap particle filter on the 4-D manoeuvring
bearings-only tracking model (BOT_Experiment_script.py:19-53), N = 2^24 particles per GPU, T = 1000 time steps,
ess_threshold = 1.0 (resample every step, :151), systematic resampling, in-kernel Philox noise, outputs =
per-step weighted mean + log-likelihood increment (the [N,T,dx] trajectory is 335 GB and is not stored).
One bench "step" = one whole filter run (all T time steps) = N*T particle-steps per GPU.  With --gpus N every
rank runs its own independent filter (trial sharding, no data-path collective): scaling = weak.
The Gaussian-sum filter number (configs[1]: UNGM, M=100 x 1024 trials, T=500) is reported beside it under
"gsf" in the same JSON line.

Keys beyond the base contract: roofline, cpu_baseline, e2e, gpu_launches, clocks, gsf.
The oracle (oracle/) is executed here only for cpu_baseline and --impl reference.
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

PF_LOG2N = int(os.environ.get("BF_BENCH_LOG2N", "24"))
# "trials": every rank runs its own N = 2^24 filter (weak scaling, no collective) -- the default;
# "particles": ONE filter of N = 2^24 particles sharded over the ranks (strong scaling; NCCL + peer gathers every step)
PF_MODE = os.environ.get("BF_BENCH_MODE", "trials")
PF_T = int(os.environ.get("BF_BENCH_T", "1000"))
GSF_M, GSF_B, GSF_T = 100, 1024, 500
BYTES_PER_PARTICLE_STEP = 84     # 4*(4*dx + 5) for dx=4, in-kernel RNG, resampling step
# DRAM bytes the fused kernel actually moves per particle-step: ncu --set full, dram__bytes_read+write of one launch /
# (N*T) (profiles/r1_m_pf_fused_T40_ncu.txt).  Below the algorithmic figure because the resampled copy of the
# particles is never materialised (the gather is fused into the next step's propagate).
MOVED_BYTES_PER_PARTICLE_STEP = 52.0
METRIC, UNIT = "particle-steps/s", "particle-steps/s"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.th.join(timeout=2)
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); smax = float(r[1])
            except ValueError:
                continue
            for k, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(k)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def _bot_case():
    from tests.parity_cases import case
    return case("bot2")


# ------------------------------------------------------------------------------------------ CPU arm
def _cpu_worker(args):
    """One process = one independent filter of the oracle (NumPy restatement of inference.py:1302-1380)."""
    seed, N, T = args
    from oracle import pf as opf
    c = _bot_case()
    _, ys = c.simulate(T, 1000 + seed)
    u = c.inputs(T)
    x0, noise, ru = c.pf_randomness(T, N, 2000 + seed, "systematic")
    t0 = time.perf_counter()
    o = opf.bootstrap_particle_filter(c.omodel, ys, u, x0, noise, ru, c.r0, c.R, 1.0, "systematic", keep_trajectory=False)
    return time.perf_counter() - t0, float(o["marginal_loglik"])


def cpu_throughput(seconds_target=12.0):
    """Oracle port on all host cores: one independent filter per core (the same trial sharding the GPU arm uses),
    N = 2^17 particles, T sized for ~seconds_target of work.  Returns (particle-steps/s, cores, sample string)."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    N = 1 << 17
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        t0 = time.perf_counter()
        pool.map(_cpu_worker, [(i, N, 2) for i in range(cores)])            # calibration + warm-up
        per = (time.perf_counter() - t0) / 2
        T = int(max(4, min(200, seconds_target / max(per, 1e-3))))
        t0 = time.perf_counter()
        pool.map(_cpu_worker, [(i, N, T) for i in range(cores)])
        wall = time.perf_counter() - t0
    return cores * N * T / wall, cores, f"{cores} independent BOT filters x N=2^17 x T={T} (NumPy oracle, one per core, {wall:.1f} s)"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    times = []
    sample = ""
    for i in range(args.warmup + args.steps):
        v, cores, sample = cpu_throughput(seconds_target=6.0 if args.steps > 1 else 12.0)
        if i >= args.warmup:
            times.append(v)
    v = float(np.mean(times))
    N = 1 << 17
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * (1 << 17) * cores / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"BPF bearings-only dx=4 dy=2, resample every step, systematic; CPU sample: {sample}"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from bayesianfiltering_b200 import inference as gf
    from bayesianfiltering_b200.engine import ParticleFilterEngine, GaussianSumFilterEngine, ShardedParticleFilterEngine
    from bayesianfiltering_b200.random import PRNGKey
    from tests.parity_cases import case

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    peak, peak_src = _peaks()
    c = _bot_case()
    N, T = 1 << PF_LOG2N, PF_T
    _, ys = c.simulate(T, 1000 + (0 if PF_MODE == 'particles' else rank))       # one filter: the same data on every rank
    u = c.inputs(T)
    ys_d, u_d = torch.from_numpy(ys).to(dev), torch.from_numpy(u).to(dev)
    sharded = PF_MODE == "particles" and world > 1
    if sharded:
        eng = ShardedParticleFilterEngine(c.host_model(), N, 1.0, "systematic", seed=1234, device=local)
        Nloc = N // world
    else:
        eng = ParticleFilterEngine(c.host_model(), N, 1, 1.0, "systematic", "streamed", 1234 + rank, local)
        Nloc = N
    outs = ("means", "log_likelihood")
    bufs = {"means": torch.empty((1, T, 4), device=dev), "log_likelihood": torch.empty((1, T), device=dev)}

    def step():
        return eng.run(ys_d, inputs=u_d, outputs=outs, out_buffers=bufs)

    for _ in range(args.warmup):
        step()
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = eng.launches()
    eng.phase_times(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    clk = clocks.stop() if rank == 0 else None
    launches = eng.launches() - l0
    phases = {k: v * 1e3 / (args.steps * T) for k, v in eng.phase_times().items()}     # us per time step
    if os.environ.get("BF_BENCH_ALLRANKS"):
        print(f"rank {rank}: phases us/step {({k: round(v, 1) for k, v in phases.items()})}", file=sys.stderr, flush=True)
    ll = bufs["log_likelihood"].double().sum().item()
    if sharded and rank != 0:
        ll = 0.0                                   # per-step scalars of a sharded filter are written by rank 0
    if not np.isfinite(ll):
        raise SystemExit(f"benchmark produced a non-finite log-likelihood ({ll})")
    ms_per_step = ms / args.steps
    value = (N if sharded else world * N) * T / (ms_per_step * 1e-3)
    # roofline: algorithmic bytes of one rank's share of a run / its device time
    achieved = Nloc * T * BYTES_PER_PARTICLE_STEP / (ms_per_step * 1e-3) / 1e9
    eng.close()

    # ---- end to end through the reference-facing entry point: host emissions in, host estimates out
    params = c.params_bpf()
    ne2e = max(1, min(args.steps, 3))
    gf.bootstrap_particle_filter(params, ys, N, PRNGKey(7), u, 1.0, resampler="systematic", outputs=outs, path="streamed",
                                 device=local, shard_particles=sharded)
    barrier()
    t0 = time.perf_counter()
    for _ in range(ne2e):
        post = gf.bootstrap_particle_filter(params, ys, N, PRNGKey(7), u, 1.0, resampler="systematic", outputs=outs,
                                            path="streamed", device=local, shard_particles=sharded)
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / ne2e
    e2e = {"value": (N if sharded else world * N) * T / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(ys.nbytes + u.nbytes),
           "d2h_bytes_per_step": int(post["means"].nbytes + post["log_likelihood_increments"].nbytes),
           "call": "inference.bootstrap_particle_filter(params, emissions[host], N, key, inputs, 1.0, "
                   "resampler='systematic', outputs=('means','log_likelihood'))", "ms_per_step": e2e_ms}

    # ---- Gaussian-sum filter, configs[1]: UNGM, M=100 x 1024 trials x T=500 (per rank)
    g = case("ungm")
    gy = np.stack([g.simulate(GSF_T, 50000 + rank * GSF_B + b)[1] for b in range(16)])
    gy = np.tile(gy, (GSF_B // 16, 1, 1)) + np.random.default_rng(rank).normal(0, 0.1, (GSF_B, GSF_T, 1)).astype(np.float32)
    gu = g.inputs(GSF_T)
    gm0 = np.stack([g.gsf_means0(GSF_M, 60000 + b) for b in range(GSF_B)])
    geng = GaussianSumFilterEngine(g.host_model(), GSF_M, GSF_B, local)
    gy_d, gu_d, gm0_d = (torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev) for a in (gy, gu, gm0))
    fields = ("weights", "means", "covariances", "predicted_means", "predicted_covariances")
    gb = {k: torch.empty((GSF_B, GSF_T, GSF_M) + ((1,) if "mean" in k else (1, 1) if "cov" in k else ()), device=dev)
          for k in fields}
    for _ in range(args.warmup):
        geng.run(gy_d, inputs=gu_d, init_means=gm0_d, outputs=fields, out_buffers=gb)
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(args.steps):
        geng.run(gy_d, inputs=gu_d, init_means=gm0_d, outputs=fields, out_buffers=gb)
    g1.record()
    barrier()
    gms = max_over_ranks(g0.elapsed_time(g1)) / args.steps
    gval = world * GSF_M * GSF_B * GSF_T / (gms * 1e-3)
    t0 = time.perf_counter()
    geng.run_host(gy, inputs=gu, init_means=gm0, outputs=fields)
    ge2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    geng.close()
    gsf = {"metric": "component-steps/s", "value": gval, "unit": "component-steps/s", "ms_per_step": gms,
           "config": {"workload": f"GSF(EKF) UNGM M={GSF_M} x {GSF_B} trials x T={GSF_T} per GPU, all five posterior fields written"},
           "roofline": {"bound": "hbm", "achieved": GSF_M * GSF_B * GSF_T * 20 / (gms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": GSF_M * GSF_B * GSF_T * 20 / (gms * 1e-3) / 1e9 / peak, "traffic": None,
                        "note": "20 B/component-step of outputs; state is register-resident"},
           "e2e": {"value": world * GSF_M * GSF_B * GSF_T / (ge2e_ms * 1e-3), "unit": "component-steps/s",
                   "h2d_bytes_per_step": int(gy.nbytes + gu.nbytes + gm0.nbytes),
                   "d2h_bytes_per_step": int(GSF_M * GSF_B * GSF_T * 20)}}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, cores, sample = cpu_throughput()
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"BPF bearings-only (BOT) dx=4 dq=2 dy=2, N=2^{PF_LOG2N} particles {'in total' if sharded else 'per GPU'}, T={T}, "
                                       "ess_threshold=1.0 (resample every step), systematic resampling, Philox noise in-kernel, "
                                       "outputs=weighted mean + log-lik per step",
                           "parallelism": (f"ONE filter, particles sharded x{world} (NCCL: weight statistics + scan offsets per step; "
                                           f"ancestors on other ranks gathered over NVLink peer memory)") if sharded else
                                          f"trial-sharded x{world} (one independent filter per GPU, no collective)",
                           "l2_policy": "inputs larger than L2 (268 MB particle state per buffer)",
                           "step": "one whole filter run = N*T particle-steps"},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": MOVED_BYTES_PER_PARTICLE_STEP * Nloc * T, "traffic_source": "ncu capture of a T=40 launch, scaled to this T",
                             "moved_bytes_per_particle_step": MOVED_BYTES_PER_PARTICLE_STEP,
                             "peak_source": peak_src, "bytes_per_particle_step": BYTES_PER_PARTICLE_STEP,
                             "bound_note": "this model is instruction-issue-bound (~475 SASS instructions per particle-step, issue slots 65% busy, DRAM 32% of peak: profiles/r1_m_pf_fused_T40_ncu.txt)",
                             "kernel": "k_pf_fused (persistent cooperative kernel, all T steps in one launch)",
                             "phase_us_per_time_step": phases},
                "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "gsf": gsf}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        print(f"note: warmup {args.warmup} < 3; timing rules ask for >= 3", file=sys.stderr)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
