#!/usr/bin/env python3
"""bench.py -- options priced per second on N B200s, with the HBM roofline and the CPU reference beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Metric (BASELINE.json): options priced per second.  Every evaluation counts: one STEP is one ROI of the
reference driver, i.e. NUM_RUNS=100 passes of the Map over the whole option set (blackscholes.c:318).

  value   whole-job rate with the SoA streams resident in HBM when the timed region starts: per step,
          NUM_RUNS real kernel launches per GPU, timed with CUDA events on the stream that launches them
          (inside libbs_gpu.so), max over ranks, K steps between barriers.
  e2e     the same metric through the public C-ABI call bs_gpu_price() with HOST buffers: every step copies
          the six input streams H2D from pinned memory, runs NUM_RUNS launches, and copies the prices D2H.
  roofline   HBM: 28 B/option (fp32) x options per launch / average launch duration, against the measured
          copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline   the UNMODIFIED reference FastFlow build (oracle/_ref/bs_ref_ff, compiled from
          /root/reference by oracle/Makefile) on this box's host cores, on a bounded sample.

Default workload = BASELINE.json configs[1]: blackscholes native, 10M options, fp32, NUM_RUNS=100 per GPU
(weak scaling: each rank prices its own 10M-option set; shards are independent, there is no collective).
Inputs (280 MB per GPU) exceed the 126 MB L2, so consecutive runs cannot be served from cache.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))  # oracle_lib (checker / CPU-baseline legs only)

NUM_RUNS = 100  # blackscholes.c:87

WORKLOADS = {
    # name: (options per GPU or total, fp_bytes, scaling, description)
    "native": (10_000_000, 4, "weak", "blackscholes native (10M options, fp32, NUM_RUNS=100) per GPU"),
    "simsmall": (4_096, 4, "weak", "blackscholes simsmall (4,096 options, fp32, NUM_RUNS=100) per GPU"),
    "native_fp64": (10_000_000, 8, "weak", "blackscholes native 10M options, fptype=double, NUM_RUNS=100 per GPU"),
    "synth1b": (1_000_000_000, 4, "strong", "synthetic 1B-option fp32 set (inputgen distribution) sharded over the GPUs, NUM_RUNS=100"),
}
# the sibling Map has its own tool with the same JSON contract: `--workload swaptions_<set>` forwards to it
SW_WORKLOADS = ("native", "simlarge", "simmedium", "simsmall")
CPU_SAMPLE_OPTIONS = 2_000_000  # bounded sample for the CPU reference: 2M options x 100 runs


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic_bytes(workload):
    """dram bytes per launch of the pricing kernel from the committed ncu capture, if one exists."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[workload]["dram_bytes_per_launch"]
    except Exception:
        return None


class ClockSampler:
    """SM clock, power and throttle reasons of one GPU sampled through NVML every ~2 ms WHILE the timed region
    runs (a Python thread; the timed calls are ctypes calls that release the GIL)."""

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.samples = []
        self._stop = False
        self._thread = None
        self.err = None

    def _loop(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.idx)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            while not self._stop:
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                try:
                    reasons = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    reasons = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                try:
                    power = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
                except Exception:
                    power = None
                self.samples.append((sm, reasons, power))
                time.sleep(0.002)
            pynvml.nvmlShutdown()
        except Exception as e:  # NVML missing: report it in the JSON instead of failing the run
            self.err = str(e)

    def start(self):
        import threading
        self._thread = threading.Thread(target=self._loop, daemon=True)
        self._thread.start()
        time.sleep(0.05)

    def stop(self):
        self._stop = True
        if self._thread is not None:
            self._thread.join(timeout=5)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: %s" % self.err]}
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                 0x80: "hw_power_brake_slowdown"}
        seen = set()
        for _, r, _ in self.samples:
            for bit, nm in names.items():
                if r & bit:
                    seen.add(nm)
        sm = sorted(x[0] for x in self.samples)
        pw = [x[2] for x in self.samples if x[2] is not None]
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": getattr(self, "max_mhz", None),
                "reasons": sorted(seen), "samples": len(sm), "power_w_max": max(pw) if pw else None, "source": "nvml"}


# ---------------------------------------------------------------------------------------------------
# CPU reference: the unmodified FastFlow build on the host cores
# ---------------------------------------------------------------------------------------------------
def host_cores():
    return len(os.sched_getaffinity(0))


def run_cpu_reference(n_options, steps, warmup, fp_bytes=4):
    """Times oracle/_ref/bs_ref_ff (or the fp64 build) on `n_options` inputgen options x NUM_RUNS, all cores.
    Returns (options_per_s, cores, kind, sample_description, per-step roi seconds)."""
    import oracle_lib  # checker/baseline leg: the one place bench.py executes oracle/
    exe_name = "bs_ref_ff" if fp_bytes == 4 else "bs_ref_ff_fp64"
    cores = host_cores()
    gen = os.path.join(ROOT, "p3arsec_b200", "bin", "bs_inputgen")
    tmp = tempfile.mkdtemp(prefix="bs_ref_")
    inp, out = os.path.join(tmp, "in.txt"), os.path.join(tmp, "out.txt")
    subprocess.run([gen, str(n_options), inp], check=True, stdout=subprocess.DEVNULL)
    rois = []
    kind = "reference"
    try:
        if oracle_lib.ref_binary(exe_name) is None:
            raise FileNotFoundError(exe_name)
        for i in range(warmup + steps):
            _, roi = oracle_lib.run_ref(exe_name, cores, inp, out, timeout=1800)
            if i >= warmup:
                rois.append(roi)
        sample = "first %d options of the workload x NUM_RUNS=%d through oracle/_ref/%s (unmodified reference FastFlow build, " \
                 "%d worker threads), ROI time as printed at blackscholes.c:781/912" % (n_options, NUM_RUNS, exe_name, cores)
    except FileNotFoundError:
        # reference binary absent: time the oracle port (OpenMP over the same body) instead
        kind = "port"
        d = oracle_lib.load(inp, fp_bytes)
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            for _ in range(NUM_RUNS):
                oracle_lib.price_map(d["sptprice"], d["strike"], d["rate"], d["volatility"], d["otime"], d["otype"], fp_bytes, cores)
            if i >= warmup:
                rois.append(time.perf_counter() - t0)
        sample = "first %d options x NUM_RUNS=%d through the oracle port (OpenMP, %d threads)" % (n_options, NUM_RUNS, cores)
    finally:
        for p in (inp, out):
            if os.path.exists(p):
                os.unlink(p)
        os.rmdir(tmp)
    rate = n_options * NUM_RUNS * len(rois) / sum(rois)
    return rate, cores, kind, sample, rois


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n_total, fp_bytes, scaling, desc = WORKLOADS[args.workload]
    # bounded sample: keep (K+W) whole-process runs of the reference binary within a few minutes
    n_sample = min(n_total, max(100_000, int(CPU_SAMPLE_OPTIONS * min(1.0, 13.0 / (args.steps + args.warmup)))))
    t0 = time.perf_counter()
    rate, cores, kind, sample, rois = run_cpu_reference(n_sample, args.steps, args.warmup, fp_bytes)
    line = {
        "impl": "reference", "metric": "options_priced_per_sec", "value": rate, "unit": "options/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(rois) / len(rois), "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": "f32" if fp_bytes == 4 else "f64", "data": "synthetic",
        "config": {"workload": desc, "num_runs": NUM_RUNS, "sample_options": n_sample,
                   "note": "CPU reference on the host cores of this box; value = sample options x NUM_RUNS / roi.time"},
        "cpu_baseline": {"value": rate, "unit": "options/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": rate, "unit": "options/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def table_soa(fp_bytes):
    """The 1000-row inputgen table as SoA arrays, obtained the way the driver obtains it: text -> loader."""
    from p3arsec_b200 import host
    gen = os.path.join(ROOT, "p3arsec_b200", "bin", "bs_inputgen")
    fd, path = tempfile.mkstemp(prefix="bs_table_", suffix=".txt")
    os.close(fd)
    try:
        subprocess.run([gen, "1000", path], check=True, stdout=subprocess.DEVNULL)
        return host.load_options(path, fp_bytes)
    finally:
        os.unlink(path)


def ours(args):
    import numpy as np
    import torch
    from p3arsec_b200 import host
    from p3arsec_b200.dist import Ranks, shard_range

    rank, local_rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this benchmark has no CPU path (use --impl reference for the CPU baseline)")
    n_cfg, fp_bytes, scaling, desc = WORKLOADS[args.workload]
    in_process_gpus = 1
    if world == 1 and args.gpus > 1:
        in_process_gpus = args.gpus  # launched without torchrun: one context drives all GPUs (one host thread each)
    numa_cpus = None
    if world > 1 and not args.no_numa_bind:
        from p3arsec_b200.dist import bind_to_gpu_locality
        numa_cpus = bind_to_gpu_locality(local_rank)  # before any staging memory is touched
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ranks = Ranks(backend="nccl", device=dev)
    n_gpus = world * in_process_gpus

    if scaling == "weak":
        n_local = n_cfg * in_process_gpus
        first_index = rank * n_cfg
    else:
        first_index, n_local = shard_range(n_cfg, world, rank)
    n_total = ranks.sum(n_local)
    devices = list(range(in_process_gpus)) if in_process_gpus > 1 else [local_rank]
    math = {"default": host.MATH_DEFAULT, "ieee": host.MATH_IEEE, "fast": host.MATH_FAST}[args.math]
    host_staging = args.workload != "synth1b"  # 1B options: device-resident only (24 GB of host staging otherwise)

    bs = host.BlackScholesGPU(n_local, fp_bytes=fp_bytes, devices=devices, math=math, host_staging=host_staging,
                              with_dgrefval=False, unroll=args.unroll, threads_per_block=args.threads, blocks_per_sm=args.blocks_per_sm)
    launch = bs.launch()

    # ---- inputs: the cyclic inputgen set.  Host copy (for e2e) = table tiled; device copy = same rows.
    if host_staging:
        tab = table_soa(fp_bytes)
        reps = -(-n_local // 1000)
        shift = first_index % 1000
        for name in ("sptprice", "strike", "rate", "volatility", "otime", "otype"):
            bs.host(name)[:] = np.tile(np.roll(tab[name], -shift), reps)[:n_local]
        bs.mark_dirty()
        bs.upload()
    else:
        bs.fill_synthetic(first_index)

    # ---- value: data resident in HBM -------------------------------------------------------------
    for _ in range(args.warmup):
        bs.run(NUM_RUNS)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    torch.cuda.synchronize()
    ranks.barrier()
    t0 = time.perf_counter()
    dev_ms, launches = 0.0, 0
    for _ in range(args.steps):
        bs.run(NUM_RUNS)
        tm = bs.timing()
        dev_ms += tm["roi_ms"]
        launches += tm["kernel_launches"]
    torch.cuda.synchronize()
    ranks.barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if rank == 0 else None
    dev_ms_max = ranks.max(dev_ms)
    wall_ms_max = ranks.max(wall_ms)
    launches_total = int(ranks.sum(launches))
    units = n_total * NUM_RUNS * args.steps
    value = units / (dev_ms_max * 1e-3)

    # parity spot check of what was just timed (checker only; not inside any timed region)
    spot = None
    if rank == 0:
        try:
            import oracle_lib
            k = min(n_local, 4096)
            ins = [bs.read_device(nm, 0, k) for nm in ("sptprice", "strike", "rate", "volatility", "otime", "otype")]
            got = bs.read_device("prices", 0, k).astype(np.float64)
            ref = oracle_lib.price_map(*ins, fp_bytes=fp_bytes).astype(np.float64)
            spot = float(np.abs(got - ref).max())
        except Exception as e:  # oracle not built on this box: report, do not fail the measurement
            spot = "unchecked: %s" % e

    # ---- e2e: host buffers -> bs_gpu_price -> host prices ------------------------------------------
    e2e = None
    if host_staging:
        for _ in range(min(args.warmup, 3)):
            bs.mark_dirty()
            bs.price(NUM_RUNS)
        torch.cuda.synchronize()
        ranks.barrier()
        t0 = time.perf_counter()
        h2d = d2h = 0
        checksum = 0.0
        for _ in range(args.steps):
            bs.mark_dirty()                      # this step's inputs are "new": forces the H2D
            bs.price(NUM_RUNS)
            tm = bs.timing()
            h2d += tm["h2d_bytes"]
            d2h += tm["d2h_bytes"]
            checksum += float(bs.prices[:: max(1, n_local // 1024)].sum())  # read the result on the host
        torch.cuda.synchronize()
        ranks.barrier()
        e2e_ms_max = ranks.max((time.perf_counter() - t0) * 1e3)
        e2e = {"value": units / (e2e_ms_max * 1e-3), "unit": "options/s",
               "h2d_bytes_per_step": int(ranks.sum(h2d) / args.steps), "d2h_bytes_per_step": int(ranks.sum(d2h) / args.steps),
               "ms_per_step": e2e_ms_max / args.steps, "checksum": checksum,
               "path": "pinned host SoA -> bs_gpu_price(): H2D, NUM_RUNS launches, D2H -> pinned host prices"}

    # ---- roofline of the pricing kernel (rank 0's GPU) --------------------------------------------
    peak, peak_src = measured_peak_gbs()
    bpo = host.bytes_per_option(fp_bytes)
    n_per_launch = n_local // in_process_gpus
    avg_launch_ms = dev_ms / (args.steps * NUM_RUNS)
    achieved = bpo * n_per_launch / (avg_launch_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic_bytes(args.workload), "peak_source": peak_src, "frac_of_nominal_8000": achieved / 8000.0,
                "kernel": "bsk::bs_map<%s, ...>" % ("float" if fp_bytes == 4 else "double"), "algorithmic_bytes_per_option": bpo,
                "options_per_launch": n_per_launch, "avg_launch_us": avg_launch_ms * 1e3}

    # ---- CPU reference beside it (rank 0, N=1 only, bounded sample) ------------------------------
    cpu = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        try:
            rate, cores, kind, sample, _ = run_cpu_reference(min(CPU_SAMPLE_OPTIONS, n_cfg), 1, 1, fp_bytes)
            cpu = {"value": rate, "unit": "options/s", "cores": cores, "kind": kind, "sample": sample}
        except Exception as e:
            cpu = {"value": None, "unit": "options/s", "cores": host_cores(), "kind": "unavailable", "sample": str(e)}

    bs.close()
    ranks.close()
    if rank != 0:
        return 0
    line = {
        "metric": "options_priced_per_sec", "value": value, "unit": "options/s", "n_gpus": n_gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "f32" if fp_bytes == 4 else "f64", "data": "synthetic",
        "config": {"workload": desc, "options_total": int(n_total), "num_runs": NUM_RUNS, "math": launch["math"],
                   "threads_per_block": launch["threads_per_block"], "blocks": launch["blocks"],
                   "parallelism": "%d independent contiguous shards, no collective" % n_gpus,
                   "numa": ("rank 0 bound to %d GPU-local CPUs" % len(numa_cpus)) if numa_cpus else "no binding",
                   "l2": "inputs+outputs per GPU = %.0f MB vs 126 MB L2 (inputs larger than L2; no flush needed)" % (bpo * n_per_launch / 1e6)
                         if bpo * n_per_launch > 126e6 else "working set fits L2: runs after the first are L2-resident",
                   "step": "one ROI = NUM_RUNS launches over the whole set (blackscholes.c:318)"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches_total, "roofline": roofline, "cpu_baseline": cpu,
        "wall_ms_per_step": wall_ms_max / args.steps, "parity_spot_max_abs": spot,
    }
    print(json.dumps(line), flush=True)
    return 0


def ensure_built():
    """Built artefacts are git-ignored; if this checkout has none yet, build them once (nvcc is part of the image)."""
    lib = os.path.join(ROOT, "p3arsec_b200", "lib", "libbs_gpu.so")
    gen = os.path.join(ROOT, "p3arsec_b200", "bin", "bs_inputgen")
    if os.path.exists(lib) and os.path.exists(gen) and os.path.exists(os.path.join(ROOT, "p3arsec_b200", "lib", "libsw_gpu.so")):
        return
    if int(os.environ.get("LOCAL_RANK", "0")) == 0:
        import __graft_entry__
        __graft_entry__.build()
    else:  # other ranks wait for rank 0's build
        for _ in range(600):
            if os.path.exists(lib) and os.path.exists(gen):
                break
            time.sleep(1.0)


def main():
    ensure_built()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS) + ["swaptions_" + w for w in SW_WORKLOADS], default="native")
    ap.add_argument("--math", choices=["default", "ieee", "fast"], default="default")
    ap.add_argument("--unroll", type=int, default=0)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--blocks-per-sm", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true", help="multi-rank runs: do not pin each rank to its GPU's NUMA node")
    args = ap.parse_args()
    if args.workload.startswith("swaptions_"):
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import sw_bench
        return sw_bench.main(["--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup), "--impl", args.impl,
                              "--workload", args.workload[len("swaptions_"):]] + (["--no-cpu-baseline"] if args.no_cpu_baseline else []))
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        return reference_arm(args)
    return ours(args)


if __name__ == "__main__":
    sys.exit(main())
