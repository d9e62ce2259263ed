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
    runs (a Python thread; the timed calls are ctypes calls that release the GIL).  start() returns once NVML is up
    and the first sample exists, so that even an 80 ms region is covered; only samples taken between start() and
    stop() are reported."""

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.samples = []
        self._stop = False
        self._thread = None
        self._armed = False
        self.err = None
        import threading
        self._ready = threading.Event()
        self._thread = threading.Thread(target=self._loop, daemon=True)
        self._thread.start()   # NVML initialisation (tens of ms) happens now, long before the timed region

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
                if self._armed:
                    self.samples.append((sm, reasons, power))
                self._ready.set()
                time.sleep(0.002 if self._armed else 0.0005)
            pynvml.nvmlShutdown()
        except Exception as e:  # NVML missing: report it in the JSON instead of failing the run
            self.err = str(e)
            self._ready.set()

    def start(self):
        self._ready.wait(timeout=10)
        self._armed = True

    def stop(self):
        self._armed = False
        self._stop = True
        if self._thread is not None:
            self._thread.join(timeout=5)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: %s" % self.err]}
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                 0x80: "hw_power_brake_slowdown"}
        seen = {}
        for _, r, _ in self.samples:
            for bit, nm in names.items():
                if r & bit:
                    seen[nm] = seen.get(nm, 0) + 1
        sm = sorted(x[0] for x in self.samples)
        pw = [x[2] for x in self.samples if x[2] is not None]
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": getattr(self, "max_mhz", None),
                "reasons": sorted(seen), "reason_samples": seen, "samples": len(sm), "power_w_max": max(pw) if pw else None, "source": "nvml"}


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
    # bounded sample: keep (K+W) whole-process runs of the reference binary within a few minutes.  A whole-process run
    # of the reference on the full native set costs ~25 s (its serial fscanf loader and fprintf writer, blackscholes.c:
    # 726-734,935-942, dominate), so the full set is used for every step only when K+W <= 4; otherwise the K steps run
    # on the sample and ONE extra run prices the full set, reported under `full_size`.
    runs = args.steps + args.warmup
    full_fits = n_total <= 10_000_000
    if full_fits and (runs <= 4 or n_total <= CPU_SAMPLE_OPTIONS):
        n_sample = n_total
    else:
        n_sample = min(n_total, max(100_000, int(CPU_SAMPLE_OPTIONS * min(1.0, 13.0 / runs))))
    t0 = time.perf_counter()
    rate, cores, kind, sample, rois = run_cpu_reference(n_sample, args.steps, args.warmup, fp_bytes)
    full = None
    if full_fits and n_sample < n_total and not args.no_full_size:
        try:
            frate, _, fkind, fsample, frois = run_cpu_reference(n_total, 1, 0, fp_bytes)
            full = {"value": frate, "unit": "options/s", "options": n_total, "roi_s": frois[0], "kind": fkind, "runs": 1,
                    "what": "ONE whole-process run of the reference on the full workload (all %d options x NUM_RUNS), same binary, same cores" % n_total}
        except Exception as e:
            full = {"value": None, "error": str(e)}
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
    if full is not None:
        line["full_size"] = full
        line["wall_s"] = time.perf_counter() - t0
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


INPUT_NAMES = ("sptprice", "strike", "rate", "volatility", "otime", "otype")


def mem_available_bytes():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) * 1024
    except Exception:
        pass
    return None


def load_inputs(bs, n_local, first_index, fp_bytes, host_staging):
    """The cyclic inputgen set, option i of the job = table row i % 1000.  With host staging the HOST buffers hold it
    (so that e2e really copies it every step) and it is uploaded once for the device-resident measurement."""
    import numpy as np
    if not host_staging:
        bs.fill_synthetic(first_index)
        return
    if n_local <= 64_000_000:
        tab = table_soa(fp_bytes)
        reps = -(-n_local // 1000)
        shift = first_index % 1000
        for name in INPUT_NAMES:
            bs.host(name)[:] = np.tile(np.roll(tab[name], -shift), reps)[:n_local]
        bs.mark_dirty()
        bs.upload()
        return
    # large sets: generated on the device (same table, same text round trip), then copied into the host staging
    # buffers so that every e2e step has real host data to send
    bs.fill_synthetic(first_index)
    step = 32_000_000
    for name in INPUT_NAMES:
        h = bs.host(name)
        for a in range(0, n_local, step):
            bs.read_device_into(name, a, h[a:min(n_local, a + step)])


def timed_resident(bs, steps, warmup, ranks, torch, sampler=None):
    """`value`: K steps of NUM_RUNS launches over device-resident streams, CUDA events inside the library, barrier +
    synchronize on both sides.  Returns this rank's (device ms, launches, wall ms) and the clock samples."""
    for _ in range(warmup):
        bs.run(NUM_RUNS)
    if sampler is not None:
        sampler.start()
    torch.cuda.synchronize()
    ranks.barrier()
    t0 = time.perf_counter()
    dev_ms, launches = 0.0, 0
    for _ in range(steps):
        bs.run(NUM_RUNS)
        tm = bs.timing()
        dev_ms += tm["roi_ms"]
        launches += tm["kernel_launches"]
    torch.cuda.synchronize()
    ranks.barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if sampler is not None else None
    return dev_ms, launches, wall_ms, clocks


def timed_e2e(bs, n_local, steps, warmup, ranks, torch, units):
    """`e2e`: K calls of bs_gpu_price() on HOST buffers -- H2D of the six streams, NUM_RUNS launches, D2H of the
    prices, result read on the host -- wall clock between barriers, max over ranks."""
    for _ in range(warmup):
        bs.mark_dirty()
        bs.price(NUM_RUNS)
    torch.cuda.synchronize()
    ranks.barrier()
    t0 = time.perf_counter()
    h2d = d2h = 0
    checksum = 0.0
    for _ in range(steps):
        bs.mark_dirty()                      # this step's inputs are "new": forces the H2D
        bs.price(NUM_RUNS)
        tm = bs.timing()
        h2d += tm["h2d_bytes"]
        d2h += tm["d2h_bytes"]
        checksum += float(bs.prices[:: max(1, n_local // 1024)].sum())  # read the result on the host
    torch.cuda.synchronize()
    ranks.barrier()
    ms_max = ranks.max((time.perf_counter() - t0) * 1e3)
    return {"value": units / (ms_max * 1e-3), "unit": "options/s",
            "h2d_bytes_per_step": int(ranks.sum(h2d) / steps), "d2h_bytes_per_step": int(ranks.sum(d2h) / steps),
            "ms_per_step": ms_max / steps, "steps": steps, "checksum": checksum}


def timed_copies(bs, ranks, torch, reps=4):
    """The copy ceiling of THIS run: the same pinned buffers, the same ranks at the same time, copies only
    (bs_gpu_upload: six streams H2D; bs_gpu_download: prices D2H), device events inside the library, max over ranks."""
    bs.mark_dirty()
    bs.upload()
    bs.download()
    torch.cuda.synchronize()
    ranks.barrier()
    h2d_ms = d2h_ms = 0.0
    h2d_b = d2h_b = 0
    for _ in range(reps):
        bs.mark_dirty()
        bs.upload()
        tm = bs.timing()
        h2d_ms += tm["h2d_ms"]
        h2d_b = tm["h2d_bytes"]
        bs.download()
        tm = bs.timing()
        d2h_ms += tm["d2h_ms"]
        d2h_b = tm["d2h_bytes"]
    torch.cuda.synchronize()
    ranks.barrier()
    h2d_ms, d2h_ms = ranks.max(h2d_ms / reps), ranks.max(d2h_ms / reps)
    return {"h2d_gbs_per_gpu": h2d_b / (h2d_ms * 1e-3) / 1e9 if h2d_ms > 0 else None,
            "d2h_gbs_per_gpu": d2h_b / (d2h_ms * 1e-3) / 1e9 if d2h_ms > 0 else None,
            "h2d_ms": h2d_ms, "d2h_ms": d2h_ms, "ranks_copying_at_once": ranks.world}


def live_ncu_traffic(n, fp_bytes):
    """dram__bytes_read + dram__bytes_write per launch of the pricing kernel, measured NOW with ncu on a small
    subprocess (three profiled launches of the same kernel on the same set; never the timed region).  None when
    ncu is missing or the counters are not permitted on this box."""
    import shutil
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None, "ncu not found"
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "-k", "regex:bs_map", "-s", "2", "-c", "3",
           "--csv", sys.executable, os.path.join(ROOT, "tools", "profile_target.py"), "--n", str(n), "--fp", str(fp_bytes), "--runs", "6"]
    try:
        cp = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    except Exception as e:
        return None, "ncu failed: %s" % e
    import csv
    import io
    rd, wr = [], []
    unit_scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for row in csv.reader(io.StringIO(cp.stdout)):
        if len(row) < 3 or "dram__bytes" not in ",".join(row):
            continue
        try:
            name = next(c for c in row if c.startswith("dram__bytes"))
            i = row.index(name)
            val = float(row[i + 2].replace(",", "")) * unit_scale.get(row[i + 1], 1.0)
        except Exception:
            continue
        (rd if "read" in name else wr).append(val)
    if not rd or not wr:
        return None, "ncu gave no dram counters (rc %d): %s" % (cp.returncode, (cp.stderr or cp.stdout)[-200:].replace("\n", " "))
    return {"read": sum(rd) / len(rd), "write": sum(wr) / len(wr), "launches_profiled": len(rd)}, "live: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum on tools/profile_target.py in this run"


def measure_workload(key, args, host, ranks, torch, rank, local_rank, world, in_process_gpus, steps, warmup, *, e2e_wanted=True,
                     run_order_wanted=False, copies_wanted=False, sampler=None):
    """One workload on the launched ranks.  Returns a dict with value, e2e (and friends), per-rank launch time."""
    from p3arsec_b200.dist import shard_range
    n_cfg, fp_bytes, scaling, desc = WORKLOADS[key]
    if scaling == "weak":
        n_local = n_cfg * in_process_gpus
        first_index = rank * n_cfg
    else:
        first_index, n_local = shard_range(n_cfg, world, rank)
    n_total = ranks.sum(n_local)
    n_gpus = world * in_process_gpus
    devices = list(range(in_process_gpus)) if in_process_gpus > 1 else [local_rank]
    math = {"default": host.MATH_DEFAULT, "ieee": host.MATH_IEEE, "fast": host.MATH_FAST, "reference": host.MATH_REFERENCE}[args.math]
    bpo = host.bytes_per_option(fp_bytes)
    # host staging = e2e possible.  Large sets need n_local x 28 B of pinned host memory per rank: only if the box has it.
    stage_bytes = n_local * bpo
    avail = mem_available_bytes()
    host_staging = e2e_wanted and (stage_bytes <= 2e9 or (avail is not None and stage_bytes * world * 1.5 < avail and stage_bytes <= 16e9))
    host_staging = bool(ranks.max(0.0 if host_staging else 1.0) == 0.0)  # all ranks or none
    geo = dict(unroll=args.unroll, threads_per_block=args.threads, blocks_per_sm=args.blocks_per_sm)
    bs = host.BlackScholesGPU(n_local, fp_bytes=fp_bytes, devices=devices, math=math, host_staging=host_staging, with_dgrefval=False, **geo)
    launch = bs.launch()
    load_inputs(bs, n_local, first_index, fp_bytes, host_staging)

    dev_ms, launches, wall_ms, clocks = timed_resident(bs, steps, warmup, ranks, torch, sampler)
    dev_ms_max = ranks.max(dev_ms)
    wall_ms_max = ranks.max(wall_ms)
    launches_total = int(ranks.sum(launches))
    units = n_total * NUM_RUNS * steps
    out = {"key": key, "desc": desc, "scaling": scaling, "fp_bytes": fp_bytes, "n_local": n_local, "n_total": int(n_total), "n_gpus": n_gpus,
           "value": units / (dev_ms_max * 1e-3), "ms_per_step": dev_ms_max / steps, "wall_ms_per_step": wall_ms_max / steps,
           "launches": launches_total, "launch": launch, "clocks": clocks, "bpo": bpo, "dev_ms_local": dev_ms, "steps": steps,
           "n_per_launch": n_local // in_process_gpus, "host_staging": host_staging}

    # parity spot check of what was just timed (checker only; not inside any timed region)
    if rank == 0:
        import numpy as np
        try:
            import oracle_lib
            k = min(n_local, 4096)
            ins = [bs.read_device(nm, 0, k) for nm in INPUT_NAMES]
            got = bs.read_device("prices", 0, k).astype(np.float64)
            ref = oracle_lib.price_map(*ins, fp_bytes=fp_bytes).astype(np.float64)
            out["parity_spot_max_abs"] = float(np.abs(got - ref).max())
        except Exception as e:  # oracle not built on this box: report, do not fail the measurement
            out["parity_spot_max_abs"] = "unchecked: %s" % e

    out["e2e"] = None
    if host_staging:
        e2e_steps = steps if stage_bytes <= 2e9 else max(3, min(steps, 5))  # multi-GB steps: fewer of them, same rule (>= 3 warm-ups)
        e2e_units = n_total * NUM_RUNS * e2e_steps
        e2e = timed_e2e(bs, n_local, e2e_steps, 3, ranks, torch, e2e_units)
        e2e["path"] = "pinned host SoA -> bs_gpu_price(): H2D, NUM_RUNS launches, D2H -> pinned host prices"
        e2e["schedule"] = "sub-shards (each > L2) pipelined behind the copies; every run of every sub-shard streams from HBM (DESIGN.md 5)" \
            if stage_bytes / max(1, in_process_gpus) >= (256 << 20) else "run order"
        if copies_wanted:
            cp = timed_copies(bs, ranks, torch)
            e2e["copy_ceiling"] = cp
            # a step cannot end before its inputs have crossed the bus: bytes in / measured concurrent H2D rate
            if cp["h2d_gbs_per_gpu"]:
                floor_ms = (e2e["h2d_bytes_per_step"] / n_gpus) / (cp["h2d_gbs_per_gpu"] * 1e9) * 1e3
                e2e["h2d_floor_ms_per_step"] = floor_ms
                e2e["frac_of_h2d_ceiling"] = floor_ms / e2e["ms_per_step"]
                # inputs in AND prices out, one after the other at the rates measured above.  With one GPU copying the two
                # directions overlap fully (63 GB/s both ways against 55 one way); with eight, the box's shared uplinks do
                # not (tools/micro/h2d_ceiling: 21 GB/s per GPU for both directions together against 24.5 in alone,
                # profiles/r02_h2d_ceiling_8gpu_box.jsonl), so this sum is the step's floor there.
                if cp["d2h_gbs_per_gpu"]:
                    both_ms = floor_ms + (e2e["d2h_bytes_per_step"] / n_gpus) / (cp["d2h_gbs_per_gpu"] * 1e9) * 1e3
                    e2e["copies_in_then_out_ms_per_step"] = both_ms
                    e2e["frac_of_copies_in_then_out"] = both_ms / e2e["ms_per_step"]
        out["e2e"] = e2e
    bs.close()

    if run_order_wanted and host_staging and stage_bytes / max(1, in_process_gpus) >= (256 << 20):
        # the same e2e with BS_GPU_FLAG_NO_SUBSHARDS: runs in the reference's order (blackscholes.c:318), copies hidden only
        # behind the first and the last run
        bs2 = host.BlackScholesGPU(n_local, fp_bytes=fp_bytes, devices=devices, math=math, host_staging=True, with_dgrefval=False,
                                   subshards=False, **geo)
        load_inputs(bs2, n_local, first_index, fp_bytes, True)
        ro = timed_e2e(bs2, n_local, steps, 3, ranks, torch, units)
        ro["schedule"] = "run order (BS_GPU_FLAG_NO_SUBSHARDS)"
        out["e2e_run_order"] = ro
        bs2.close()
    return out


def probe_launch_us(key, args, host, local_rank, steps):
    """The traffic-only probe (same seven streams, same geometry, five adds instead of the pricing) on the same GPU,
    same run: the ceiling of THIS access pattern on THIS board."""
    n_cfg, fp_bytes, scaling, desc = WORKLOADS[key]
    with host.BlackScholesGPU(n_cfg, fp_bytes=fp_bytes, devices=[local_rank], host_staging=False, with_dgrefval=False, variant=2,
                              unroll=args.unroll, threads_per_block=args.threads, blocks_per_sm=args.blocks_per_sm) as pb:
        pb.fill_synthetic(0)
        for _ in range(3):
            pb.run(NUM_RUNS)
        ms = 0.0
        for _ in range(steps):
            pb.run(NUM_RUNS)
            ms += pb.timing()["roi_ms"]
    return ms / (steps * NUM_RUNS) * 1e3


def inproc_record(n_dev, args, host, steps):
    """north_star item 3 as ONE process (rank 0, after the ranks of the torchrun job have let go of their GPUs):
    bs_gpu_init over all `n_dev` devices -- one host thread, context and stream per device, contiguous shards, gather
    = each device's D2H into its slice of the single pinned prices array (skepu2 map_cu.inl:141-298's role).
    (1) the native set through bs_gpu_price(): prices must be BIT-EQUAL to the single-device prices;
    (2) the 1B-option set, device-resident: the in-process twin of the headline number."""
    import numpy as np
    n = WORKLOADS["native"][0]
    tab = table_soa(4)
    reps = -(-n // 1000)
    rec = {"shards": n_dev, "steps": steps}
    with host.BlackScholesGPU(n, devices=[0], with_dgrefval=False) as one:
        for name in INPUT_NAMES:
            one.host(name)[:] = np.tile(tab[name], reps)[:n]
        one.mark_dirty()
        one.price(NUM_RUNS)
        single = one.prices.copy()
    with host.BlackScholesGPU(n, devices=list(range(n_dev)), with_dgrefval=False) as many:
        for name in INPUT_NAMES:
            many.host(name)[:] = np.tile(tab[name], reps)[:n]
        for _ in range(3):
            many.mark_dirty()
            many.price(NUM_RUNS)
        t0 = time.perf_counter()
        for _ in range(steps):
            many.mark_dirty()
            many.price(NUM_RUNS)
            _ = float(many.prices[::9973].sum())
        dt = time.perf_counter() - t0
        rec["shard_layout"] = many.shards()
        sharded = many.prices.copy()
    rec["bit_equal_to_single_device"] = bool(sharded.tobytes() == single.tobytes())
    if not rec["bit_equal_to_single_device"]:
        raise SystemExit("bench.py: in-process %d-device prices differ from the single-device prices" % n_dev)
    rec["native_e2e"] = {"value": n * NUM_RUNS * steps / dt, "unit": "options/s", "ms_per_step": dt / steps * 1e3,
                         "what": "native 10M options (strong: %d options per device) through bs_gpu_price() with host buffers" % (n // n_dev)}
    nb = WORKLOADS["synth1b"][0]
    with host.BlackScholesGPU(nb, devices=list(range(n_dev)), host_staging=False, with_dgrefval=False) as big:
        big.fill_synthetic(0)
        for _ in range(3):
            big.run(NUM_RUNS)
        ms, launches = 0.0, 0
        for _ in range(steps):
            big.run(NUM_RUNS)
            tm = big.timing()
            ms += tm["roi_ms"]          # max over the devices, CUDA events on each device's stream
            launches += tm["kernel_launches"]
        got = big.read_device("prices", nb - 1000, 1000)
    rec["value"] = nb * NUM_RUNS * steps / (ms * 1e-3)
    rec["unit"] = "options/s"
    rec["ms_per_step"] = ms / steps
    rec["gpu_launches"] = int(launches)
    rec["what"] = "synthetic 1B-option set sharded over %d devices by ONE context, device-resident, NUM_RUNS=100" % n_dev
    rec["tail_matches_single_device"] = bool(np.array_equal(got, np.roll(single[:1000], -((nb - 1000) % 1000))))
    # the same 1B-option set on ONE device of this box, same process, same minute: the N = 1 anchor of the strong-scaling curve
    with host.BlackScholesGPU(nb, devices=[0], host_staging=False, with_dgrefval=False) as one_big:
        one_big.fill_synthetic(0)
        for _ in range(3):
            one_big.run(NUM_RUNS)
        ms1 = 0.0
        for _ in range(3):
            one_big.run(NUM_RUNS)
            ms1 += one_big.timing()["roi_ms"]
    rec["single_gpu_1b"] = {"value": nb * NUM_RUNS * 3 / (ms1 * 1e-3), "unit": "options/s", "ms_per_step": ms1 / 3, "steps": 3,
                            "what": "the same 1B-option set, device-resident, on device 0 alone (this box, this process)"}
    return rec


def ours(args):
    import torch
    from p3arsec_b200 import host
    from p3arsec_b200.dist import Ranks

    rank, local_rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this benchmark has no CPU path (use --impl reference for the CPU baseline)")
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
    key = args.workload
    single = n_gpus == 1

    sampler = ClockSampler(local_rank) if rank == 0 else None
    head = measure_workload(key, args, host, ranks, torch, rank, local_rank, world, in_process_gpus, args.steps, args.warmup,
                            run_order_wanted=True, copies_wanted=True, sampler=sampler)
    n_cfg, fp_bytes, scaling, desc = WORKLOADS[key]

    # ---- the other half of north_star's metric, in the same run -----------------------------------------------
    extra = {}
    if not args.headline_only:
        if key == "native" and single:
            # N = 1 anchor of the 1B-option strong-scaling curve (device-resident; a step lasts ~0.43 s)
            s1 = measure_workload("synth1b", args, host, ranks, torch, rank, local_rank, world, in_process_gpus, max(3, min(args.steps, 5)), 3,
                                  e2e_wanted=False)
            extra["strong_1b"] = {"value": s1["value"], "unit": "options/s", "ms_per_step": s1["ms_per_step"], "steps": s1["steps"],
                                  "workload": s1["desc"], "hbm_gbs": s1["bpo"] * s1["n_per_launch"] / (s1["ms_per_step"] / NUM_RUNS * 1e-3) / 1e9,
                                  "launch": s1["launch"]}
        if key == "synth1b" and not single:
            # the weak-scaling native rate (10M options per GPU), which BENCH measures at N = 1
            w = measure_workload("native", args, host, ranks, torch, rank, local_rank, world, in_process_gpus, args.steps, args.warmup,
                                 run_order_wanted=True, copies_wanted=True)
            extra["native_weak"] = {"value": w["value"], "unit": "options/s", "ms_per_step": w["ms_per_step"], "workload": w["desc"],
                                    "e2e": w["e2e"], "e2e_run_order": w.get("e2e_run_order"),
                                    "hbm_gbs_per_gpu": w["bpo"] * w["n_per_launch"] / (w["ms_per_step"] / NUM_RUNS * 1e-3) / 1e9}

    # ---- roofline of the pricing kernel (rank 0's GPU) --------------------------------------------
    peak, peak_src = measured_peak_gbs()
    bpo = head["bpo"]
    avg_launch_ms = head["dev_ms_local"] / (args.steps * NUM_RUNS)
    achieved = bpo * head["n_per_launch"] / (avg_launch_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic_bytes(key), "traffic_source": "committed ncu capture (profiles/ncu_traffic.json)",
                "peak_source": peak_src, "frac_of_nominal_8000": achieved / 8000.0,
                "kernel": "bsk::bs_map<%s, ...>" % ("float" if fp_bytes == 4 else "double"), "algorithmic_bytes_per_option": bpo,
                "options_per_launch": head["n_per_launch"], "avg_launch_us": avg_launch_ms * 1e3}
    cpu = None
    if rank == 0 and single and not args.headline_only:
        # second denominator: the traffic-only probe on the same board in the same run
        try:
            probe_us = probe_launch_us(key, args, host, local_rank, max(3, min(args.steps, 5)))
            roofline["probe_gbs"] = bpo * n_cfg / (probe_us * 1e-6) / 1e9
            roofline["probe_launch_us"] = probe_us
            roofline["frac_of_probe"] = achieved / roofline["probe_gbs"]
        except Exception as e:
            roofline["frac_of_probe"] = None
            roofline["probe_error"] = str(e)
        if not args.no_ncu:
            live, src = live_ncu_traffic(n_cfg, fp_bytes)
            if live is not None:
                roofline["traffic"] = live["read"] + live["write"]
                roofline["traffic_read"] = live["read"]
                roofline["traffic_write"] = live["write"]
                roofline["traffic_over_algorithmic"] = roofline["traffic"] / (bpo * n_cfg)
            roofline["traffic_source"] = src if live is not None else "committed ncu capture (profiles/ncu_traffic.json); live: " + src
    # ---- CPU reference beside it (rank 0, N=1 only, bounded sample) ------------------------------
    if rank == 0 and single and not args.no_cpu_baseline:
        try:
            rate, cores, kind, sample, _ = run_cpu_reference(min(CPU_SAMPLE_OPTIONS, n_cfg), 1, 1, fp_bytes)
            cpu = {"value": rate, "unit": "options/s", "cores": cores, "kind": kind, "sample": sample}
        except Exception as e:
            cpu = {"value": None, "unit": "options/s", "cores": host_cores(), "kind": "unavailable", "sample": str(e)}

    # ---- BASELINE.json's other single-GPU configs, measured in this same run (N = 1, default workload only) ----
    other = None
    if single and key == "native" and not args.headline_only:
        other = {}
        for ok_key in ("simsmall", "native_fp64"):
            try:
                w = measure_workload(ok_key, args, host, ranks, torch, rank, local_rank, world, in_process_gpus, max(3, min(args.steps, 10)), 3)
                gbs = w["bpo"] * w["n_per_launch"] / (w["ms_per_step"] / NUM_RUNS * 1e-3) / 1e9
                other[ok_key] = {"workload": w["desc"], "value": w["value"], "unit": "options/s", "ms_per_step": w["ms_per_step"], "steps": w["steps"],
                                 "us_per_launch": w["ms_per_step"] / NUM_RUNS * 1e3, "hbm_gbs": gbs, "frac_of_measured_peak": gbs / peak,
                                 "e2e": w["e2e"], "launch": w["launch"], "parity_spot_max_abs": w.get("parity_spot_max_abs"),
                                 "dtype": "f32" if w["fp_bytes"] == 4 else "f64"}
                if ok_key == "simsmall" and not args.no_cpu_baseline:
                    rate, cores, kind, sample, _ = run_cpu_reference(WORKLOADS["simsmall"][0], 3, 1, 4)
                    other[ok_key]["cpu_baseline"] = {"value": rate, "unit": "options/s", "cores": cores, "kind": kind, "sample": sample}
            except Exception as e:
                other[ok_key] = {"error": str(e)}

    ranks.close()
    if rank != 0:
        return 0
    if other is not None:
        # the sibling Map, measured by its own tool with the same contract: a child process, after
        # this process has closed its contexts, so that its line is part of this driver-run record too
        try:
            cp = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sw_bench.py"), "--workload", "native", "--steps", str(max(3, min(args.steps, 10))),
                                 "--warmup", "3", "--no-cpu-baseline"], capture_output=True, text=True, timeout=600)
            sw = json.loads([l for l in cp.stdout.splitlines() if l.startswith("{")][-1])
            other["swaptions_native"] = {k: sw.get(k) for k in ("metric", "value", "unit", "ms_per_step", "steps", "dtype", "config", "e2e", "roofline",
                                                                 "gpu_launches", "parity_spot_max_rel", "clocks")}
        except Exception as e:
            other["swaptions_native"] = {"error": str(e)[:300]}
    inproc = None
    if world > 1 and not args.headline_only:
        time.sleep(2.0)  # the other ranks are exiting: let their contexts go before GPUs 1..N-1 are used from here
        inproc = inproc_record(world, args, host, max(3, min(args.steps, 5)))
    # ---- BASELINE.json configs[4]: end-to-end native file -> prices file through the drop-in driver, NUM_RUNS = 1,
    # <nthreads> = the GPUs of this run (the driver itself decides how many of them the work is worth) ----
    e2e_file = None
    if not args.headline_only and key in ("native", "synth1b"):
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import e2e_file_bench
            e2e_file = e2e_file_bench.measure(gpus=n_gpus, reps=2, ref_sample=300_000)
        except BaseException as e:  # SystemExit from a failed child included: report, keep the line
            e2e_file = {"error": str(e)[:500]}
    line = {
        "metric": "options_priced_per_sec", "value": head["value"], "unit": "options/s", "n_gpus": n_gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "f32" if fp_bytes == 4 else "f64", "data": "synthetic",
        "config": {"workload": desc, "options_total": head["n_total"], "num_runs": NUM_RUNS, "math": head["launch"]["math"],
                   "threads_per_block": head["launch"]["threads_per_block"], "blocks": head["launch"]["blocks"],
                   "parallelism": "%d independent contiguous shards, no collective" % n_gpus,
                   "numa": ("rank 0 bound to %d GPU-local CPUs" % len(numa_cpus)) if numa_cpus else "no binding",
                   "l2": "inputs+outputs per GPU = %.0f MB vs 126 MB L2 (inputs larger than L2; no flush needed)" % (bpo * head["n_per_launch"] / 1e6)
                         if bpo * head["n_per_launch"] > 126e6 else "working set fits L2: runs after the first are L2-resident",
                   "step": "one ROI = NUM_RUNS launches over the whole set (blackscholes.c:318)"},
        "clocks": head["clocks"], "e2e": head["e2e"], "gpu_launches": head["launches"], "roofline": roofline, "cpu_baseline": cpu,
        "wall_ms_per_step": head["wall_ms_per_step"], "parity_spot_max_abs": head.get("parity_spot_max_abs"),
    }
    if head.get("e2e_run_order"):
        line["e2e_run_order"] = head["e2e_run_order"]
    line.update(extra)
    if inproc is not None:
        line["inproc"] = inproc
    if other is not None:
        line["other_configs"] = other
    if e2e_file is not None:
        line["e2e_file"] = e2e_file
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
    ap.add_argument("--workload", choices=sorted(WORKLOADS) + ["swaptions_" + w for w in SW_WORKLOADS], default=None,
                    help="default: native at --gpus 1 (BASELINE.json configs[1]); synth1b (configs[3], the 1B-option strong-scaling "
                         "set north_star names for 1/2/4/8 GPUs) at --gpus > 1, with the weak-scaling native rate under `native_weak`")
    ap.add_argument("--math", choices=["default", "ieee", "fast", "reference"], default="default")
    ap.add_argument("--headline-only", action="store_true", help="skip the second workload, the probe, ncu and the in-process pass")
    ap.add_argument("--no-ncu", action="store_true", help="do not measure roofline.traffic live with ncu")
    ap.add_argument("--no-full-size", action="store_true", help="reference arm: skip the one extra run on the full workload")
    ap.add_argument("--unroll", type=int, default=0)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--blocks-per-sm", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true", help="multi-rank runs: do not pin each rank to its GPU's NUMA node")
    args = ap.parse_args()
    if args.workload is None:
        args.workload = "native" if args.gpus == 1 else "synth1b"
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
