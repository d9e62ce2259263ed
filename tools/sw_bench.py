#!/usr/bin/env python3
"""sw_bench.py -- Monte-Carlo trials simulated per second by the swaptions Map on N B200s,
with the FP64-pipe roofline and the CPU reference beside it.  Same JSON contract as bench.py, which forwards its
`--workload swaptions_*` runs here.

    python tools/sw_bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload native|simlarge|...]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --workload swaptions_native --gpus N --steps K --warmup W

One STEP = one ROI of the reference driver (HJM_Securities.cpp:301-345): the Map over the whole portfolio, every
swaption simulated with -sm trials.  Unit of work = one trial (one 11-point x 3-factor HJM path + discounting).

  value     trials / device time of the kernels (CUDA events on the launching stream inside libsw_gpu.so), max over
            ranks.  The per-swaption parameter records (~1.2 KB each) are resident; nothing else is input.
  e2e       the same through the public C-ABI call sw_gpu_price() with HOST arrays: host-side preparation of the
            parameter records, H2D, kernels, D2H of mean / std error, wall clock.
  roofline  this kernel is bound by the FP64 pipe, not by HBM (a trial reads no global memory): achieved = FP64-pipe
            thread-instructions per second (instructions per trial from the committed ncu capture x trials/s), peak =
            SMs x 64 lanes x SM clock.  bench.py's contract names only "hbm" and "tensor"; "fp64" is stated as such.
  cpu_baseline  oracle/_ref/sw_ref_ff (the reference's HJM_Securities.cpp + HJM_Swaption_Blocking.cpp compiled
            unmodified; the PARSEC-owned leaves they call are restated -- oracle/sw_absent/) on all host cores, on a
            bounded sample (same portfolio, fewer trials per swaption).
Strong scaling: the portfolio is fixed (PARSEC native: 128 swaptions x 1,000,000 trials) and its swaptions are split
contiguously over the GPUs; no collective on the data path.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))  # sw_oracle_lib (checker / CPU-baseline legs only)

# PARSEC input sets (parsec-3.0 pkgs/apps/swaptions/parsec/*.runconf; absent from the overlay, recalled)
WORKLOADS = {
    "native": (128, 1_000_000, "swaptions native (-ns 128 -sm 1000000)"),
    "simlarge": (64, 40_000, "swaptions simlarge (-ns 64 -sm 40000)"),
    "simmedium": (32, 20_000, "swaptions simmedium (-ns 32 -sm 20000)"),
    "simsmall": (16, 10_000, "swaptions simsmall (-ns 16 -sm 10000)"),
}
CPU_SAMPLE_TRIALS = 20_000  # per swaption, for the CPU reference
FP64_LANES_PER_SM = 64


def ncu_counts():
    """profiles/sw_ncu_counts.json: FP64-pipe thread-instructions per trial of each kernel flavour (from the committed
    ncu captures) and the DFMA rate a pure-DFMA microbenchmark reaches on this pool's B200s."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "sw_ncu_counts.json")))
    except Exception:
        return {}


def run_cpu_reference(ns, trials, steps, warmup):
    import sw_oracle_lib as so
    cores = len(os.sched_getaffinity(0))
    nt = min(cores, ns)
    rois = []
    if so.have_ref("sw_ref_ff"):
        kind = "reference"
        for i in range(warmup + steps):
            _, _, roi = so.run_ref(ns, trials, nt, None, "sw_ref_ff")
            if i >= warmup:
                rois.append(roi)
        sample = "-ns %d -sm %d -nt %d through oracle/_ref/sw_ref_ff (reference HJM_Securities.cpp + HJM_Swaption_Blocking.cpp " \
                 "unmodified, FastFlow build; PARSEC leaf routines restated), ROI time as printed at HJM_Securities.cpp:301/343" % (ns, trials, nt)
    else:
        kind = "port"
        seed, p, y, f = so.portfolio(ns)
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            so.price_map(p, y, f, seed, trials, nthreads=nt)
            if i >= warmup:
                rois.append(time.perf_counter() - t0)
        sample = "-ns %d -sm %d through the oracle port (OpenMP, %d threads)" % (ns, trials, nt)
    sims = ns * ((trials + 15) // 16) * 16
    return sims * len(rois) / sum(rois), nt, kind, sample, rois


def reference_arm(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    ns, trials, desc = WORKLOADS[args.workload]
    sample_trials = min(trials, max(1000, int(CPU_SAMPLE_TRIALS * min(1.0, 13.0 / (args.steps + args.warmup)))))
    t0 = time.perf_counter()
    rate, cores, kind, sample, rois = run_cpu_reference(ns, sample_trials, args.steps, args.warmup)
    line = {"impl": "reference", "metric": "swaption_trials_per_sec", "value": rate, "unit": "trials/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(rois) / len(rois), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "sample_trials_per_swaption": sample_trials,
                       "note": "CPU reference on the host cores of this box; value = simulated trials / roi.time"},
            "cpu_baseline": {"value": rate, "unit": "trials/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": rate, "unit": "trials/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t0}
    print(json.dumps(line), flush=True)
    return 0


def ours(args):
    import numpy as np
    import torch
    import bench as main_bench  # ClockSampler
    from p3arsec_b200 import swaptions as sw
    from p3arsec_b200.dist import Ranks, shard_range

    rank, local_rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("sw_bench.py: no CUDA device; there is no CPU path (use --impl reference for the CPU baseline)")
    ns, trials, desc = WORKLOADS[args.workload]
    if args.trials:
        trials = args.trials
    in_process_gpus = args.gpus if (world == 1 and args.gpus > 1) else 1
    torch.cuda.set_device(local_rank)
    ranks = Ranks(backend="nccl", device=torch.device("cuda", local_rank))
    n_gpus = world * in_process_gpus
    flags = {"fast": 0, "lean": sw.FLAG_LEAN, "ieee": sw.FLAG_IEEE}[args.mode] | (sw.FLAG_BATCHED if args.batched else 0)
    # which kernel the library chose is read back from its launch counts after the timed region (below)

    seed, p, y, f = sw.make_portfolio(ns)
    first, count = shard_range(ns, world, rank)
    if count == 0:
        raise SystemExit("more ranks than swaptions")
    # launched without torchrun: one context drives all GPUs; one rank per GPU: the context sits on the rank's device
    ctx = sw.SwaptionsGPU(count, num_gpus=in_process_gpus) if world == 1 else sw.SwaptionsGPU(count, devices=[local_rank])
    if args.ctas_per_sm or args.tpt:
        ctx.set_geometry(args.ctas_per_sm, args.tpt)
    ps, ys, fs = p[first:first + count], y[first:first + count], f[first:first + count]
    local_seed = seed + first  # swaption i of the portfolio uses swaption_seed + i (HJM_Securities.cpp:319)

    def step():
        return ctx.price(ps, ys, fs, local_seed, trials, sw.BLOCK_SIZE, flags)

    for _ in range(args.warmup):
        step()
    sampler = main_bench.ClockSampler(local_rank) if rank == 0 else None
    if rank == 0:
        sampler.start()
    torch.cuda.synchronize()
    ranks.barrier()
    t0 = time.perf_counter()
    dev_ms = 0.0
    launches = h2d = d2h = 0
    sims_local = 0
    checksum = 0.0
    for _ in range(args.steps):
        mean, err = step()
        tm = ctx.timing()
        dev_ms += tm["roi_ms"]
        launches += tm["kernel_launches"]
        h2d += tm["h2d_bytes"]
        d2h += tm["d2h_bytes"]
        sims_local += tm["trials_simulated"]
        checksum += float(mean.sum())
    torch.cuda.synchronize()
    ranks.barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if rank == 0 else None
    dev_ms_max, wall_ms_max = ranks.max(dev_ms), ranks.max(wall_ms)
    sims_total = ranks.sum(sims_local)
    launches_total = int(ranks.sum(launches))  # ranks may hold unequal shards: count, do not extrapolate
    # one launch per swaption (sw_sim_one) shows as count + 1 launches per device and ROI; the batched kernel as 2
    one_per_launch = launches // args.steps // max(in_process_gpus, 1) > 2
    value = sims_total / (dev_ms_max * 1e-3)
    e2e = {"value": sims_total / (wall_ms_max * 1e-3), "unit": "trials/s", "h2d_bytes_per_step": int(ranks.sum(h2d) / args.steps),
           "d2h_bytes_per_step": int(ranks.sum(d2h) / args.steps), "ms_per_step": wall_ms_max / args.steps, "checksum": checksum,
           "path": "host parm arrays -> sw_gpu_price(): host preparation, H2D, 2 kernels, D2H -> host mean/std error (wall clock)"}

    # parity spot check of what was just timed (checker only; outside every timed region)
    spot = None
    if rank == 0:
        try:
            import sw_oracle_lib as so
            k = min(count, 4)
            spot_trials = min(trials, 20000)
            gm, ge = ctx.price(ps[:k], ys[:k], fs[:k], local_seed, spot_trials, sw.BLOCK_SIZE, flags)
            om, oe = so.price_map(ps[:k], ys[:k], fs[:k], local_seed, spot_trials)
            spot = float(np.nanmax(np.abs(gm - om) / np.maximum(np.abs(om), 1e-300) * (om != 0)))
        except Exception as e:
            spot = "unchecked: %s" % e

    # roofline: FP64 pipe
    counts = ncu_counts()
    ipt = counts.get(args.mode)
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965
    props = torch.cuda.get_device_properties(local_rank)
    nominal = props.multi_processor_count * FP64_LANES_PER_SM * sm_mhz * 1e6 / 1e12  # T FP64-pipe thread-instructions / s
    peak = counts.get("fp64_peak_measured_tinst_per_s") or nominal
    per_gpu_rate = (sims_local / max(in_process_gpus, 1)) / (dev_ms * 1e-3)
    achieved = (ipt * per_gpu_rate / 1e12) if ipt else None
    roofline = {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "T fp64-pipe thread-instructions/s",
                "frac": (achieved / peak) if achieved else None, "traffic": None,
                "peak_source": counts.get("fp64_peak_source") or "nominal",
                "nominal_peak": nominal, "frac_of_nominal": (achieved / nominal) if achieved else None,
                "nominal_source": "%d SMs x %d FP64 lanes x %d MHz (clock sampled during the run)" % (props.multi_processor_count, FP64_LANES_PER_SM, sm_mhz),
                "kernel": ("swk::sw_sim_one<%s>" if one_per_launch else "swk::sw_sim_fast<%s>") % ("true" if args.mode == "lean" else "false") if args.mode != "ieee" else "swk::sw_sim_generic",
                "fp64_pipe_instructions_per_trial": ipt, "trials_per_roi_per_gpu": int(sims_local / args.steps / max(in_process_gpus, 1)),
                "roi_us": dev_ms / args.steps * 1e3, "launches_per_roi_per_gpu": launches // args.steps // max(in_process_gpus, 1),
                "launch_note": "one sw_sim_one launch per swaption, rotating over 4 streams so that their tails overlap, + sw_finalize; the ROI is timed "
                               "with CUDA events around all of them" if one_per_launch else "one simulation launch + sw_finalize per ROI",
                "note": "instructions per trial come from profiles/sw_ncu_counts.json (ncu smsp__inst_executed_pipe_fp64 x 32 / trials); "
                        "this figure is PIPE OCCUPANCY (executed instructions), see `algorithmic` for the roofline on the reference's own operation count"}
    # The algorithmic roofline: floating-point operations the REFERENCE performs per trial (HJM_Swaption_Blocking.cpp:156-204
    # and the leaves it calls, iN = 11, iFactors = 3, counted in DESIGN.md 9.5: 30 draws + 30 CumNormalInv + 55 path entries +
    # 2 discount-factor passes + payoff = ~1295 multiplies/adds/divides + 19 exp + ~10 log, each transcendental counted as
    # ONE operation as in SURVEY.md 8d) against the measured DFMA peak counted as two operations per lane and clock.
    algo_flops = counts.get("algorithmic_flops_per_trial_full", 1325.0)
    if args.mode == "fast":
        a_ach = algo_flops * per_gpu_rate / 1e12
        a_peak = 2.0 * peak
        roofline["algorithmic"] = {"achieved": a_ach, "peak": a_peak, "unit": "TFLOP/s (fp64)", "frac": a_ach / a_peak,
                                   "flops_per_trial": algo_flops,
                                   "what": "reference operation count per trial x trials/s over 2 x the measured DFMA rate"}

    cpu = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        try:
            rate, cores, kind, sample, _ = run_cpu_reference(ns, min(trials, CPU_SAMPLE_TRIALS), 1, 1)
            cpu = {"value": rate, "unit": "trials/s", "cores": cores, "kind": kind, "sample": sample}
        except Exception as e:
            cpu = {"value": None, "unit": "trials/s", "cores": len(os.sched_getaffinity(0)), "kind": "unavailable", "sample": str(e)}
    ctx.close()
    ranks.close()
    if rank != 0:
        return 0
    line = {"metric": "swaption_trials_per_sec", "value": value, "unit": "trials/s", "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "swaptions": ns, "trials_per_swaption": trials, "mode": args.mode, "block_size": sw.BLOCK_SIZE,
                       "launches": "one kernel per swaption (tables in the constant bank) + finalize" if one_per_launch else "one kernel for the portfolio + finalize",
                       "parallelism": "%d contiguous shards of the portfolio, no collective" % n_gpus,
                       "l2": "not applicable: a trial reads no global memory (per-swaption parameters sit in shared memory), so there is nothing to flush",
                       "step": "one ROI = the Map over the whole portfolio (HJM_Securities.cpp:311-323)"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches_total, "roofline": roofline,
            "cpu_baseline": cpu, "wall_ms_per_step": wall_ms_max / args.steps, "parity_spot_max_rel": spot}
    print(json.dumps(line), flush=True)
    return 0


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="native")
    ap.add_argument("--mode", choices=["fast", "lean", "ieee"], default="fast")
    ap.add_argument("--trials", type=int, default=0)
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--tpt", type=int, default=0)
    ap.add_argument("--batched", action="store_true", help="SW_GPU_FLAG_BATCHED: one kernel launch for the whole portfolio")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args(argv)
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return reference_arm(args)
    return ours(args)


if __name__ == "__main__":
    sys.exit(main())
