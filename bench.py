#!/usr/bin/env python
"""bench.py -- throughput of the negacyclic NTT hot path on B200 (one process per GPU).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json configs[1]): batched forward+inverse NTT, N = 2^14, 49-bit q = 0x1fffffc800001,
4096 polynomials per GPU, synthetic uniform coefficients (splitmix64 % q).  One step = forward transform
of the whole batch followed by the inverse transform of the whole batch (2 * batch single-direction NTTs
per GPU).  The batch is 512 MiB per GPU, four times the 126 MB L2, so every step streams from HBM.

Printed JSON (rank 0, one line): metric/value = whole-job single-direction NTTs per second over all GPUs,
timed with CUDA events on the launch stream, max over ranks; `roofline` = forward chunk kernel against the
measured HBM copy bandwidth (algorithmic bytes 2*N*8 per transform); `cpu_baseline` = the reference's own
CPU code (oracle/_ref) on this box's host cores; `e2e` = same metric through the host-buffer C-ABI calls
(ntt_b200_fwd_batch_host / ntt_b200_inv_batch_host) with pinned host memory, copies inside the timing.

--impl reference times the reference's CPU implementation (oracle/_ref, all host threads) on the same
config and prints the same line with "impl": "reference".
"""
import argparse
import importlib
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "optimized-number-theoretic-transform-implementations_b200"

LOGN = 14
Q49 = 0x1FFFFFC800001
PSI = {13: 94912374482, 14: 20456969886, 16: 3471868370}  # smallest primitive 2N-th roots
BATCH_PER_GPU = 4096
METRIC = "fwd+inv NTTs/s (N=2^14, 49-bit q, batched)"
UNIT = "NTT/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="polynomials per GPU")
    ap.add_argument("--logn", type=int, default=LOGN)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=3.0, help="target seconds per CPU baseline leg")
    return ap.parse_args()


def workload_config(args, extra=None):
    cfg = {
        "workload": "batched forward+inverse negacyclic NTT, N=2^%d, 49-bit q=%#x, batch %d per GPU "
                    "(BASELINE configs[1])" % (args.logn, Q49, args.batch),
        "N": 1 << args.logn, "q": Q49, "batch_per_gpu": args.batch,
        "step": "forward then inverse transform of the batch; value counts single-direction transforms",
        "cache": "inputs larger than L2 (batch is %d MiB per GPU)" % ((args.batch << args.logn) * 8 >> 20),
        "parallelism": "polynomials sharded across GPUs, no data-path collective",
    }
    if extra:
        cfg.update(extra)
    return cfg


def psi_for(logn, ntt=None):
    if logn in PSI:
        return PSI[logn]
    return ntt.min_primitive_root(1 << logn, Q49)


# ---- clocks ---------------------------------------------------------------------------------------------

class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU with NVML while the timed region runs."""

    REASONS = {
        0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
        0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting",
    }

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.002)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
        s = sorted(self.samples)
        med = s[len(s) // 2] if s else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ---- reference CPU arm ---------------------------------------------------------------------------------------

class CpuReference:
    """The reference's own CPU code (oracle/_ref) timed with one polynomial per thread on all host cores,
    following the reference's methodology: warm-up calls, then a timed loop that feeds each output back as
    the next input (tests/measurements.h:38-75)."""

    FWD = ["fwd_ntt_radix4_avx512_ifma", "fwd_ntt_r2_16_avx512_ifma", "fwd_ntt_radix4x4", "fwd_ntt_ref_harvey"]
    INV = ["inv_ntt_seal", "inv_ntt_ref_harvey", "inv_ntt_radix4"]

    def __init__(self, logn, threads=None):
        from oracle.pyoracle import Oracle, Reference
        self.ref, orc = Reference(), Oracle()
        self.available = self.ref.available
        self.logn = logn
        self.threads = threads or os.cpu_count() or 1
        self.psi = PSI[logn] if logn in PSI else orc.min_root(1 << logn, Q49)
        self.psi_inv, self.n_inv = orc.invmod(self.psi, Q49), orc.invmod(1 << logn, Q49)
        self.rates = {}
        self.calls = {}

    def time_variant(self, variant, seconds):
        """Aggregate transforms/s of one variant over a sample of about `seconds`."""
        b = lambda calls: self.ref.bench(variant, self.logn, Q49, self.psi, self.psi_inv, self.n_inv,
                                         self.threads, calls)
        probe = b(20)
        if probe <= 0:
            return None
        calls = max(20, int(probe / self.threads * seconds))
        self.rates[variant], self.calls[variant] = b(calls), calls
        return self.rates[variant]

    def survey(self, seconds):
        """Times every candidate once and picks the north star's forward (radix-4 AVX512-IFMA where the
        host supports it, otherwise radix4x4) and the fastest inverse (no SIMD inverse exists)."""
        for v in self.FWD + self.INV:
            if v.endswith("ifma") and not self.ref.ifma:
                continue
            self.time_variant(v, seconds)
        self.fwd_name = "fwd_ntt_radix4_avx512_ifma" if "fwd_ntt_radix4_avx512_ifma" in self.rates \
            else "fwd_ntt_radix4x4"
        self.inv_name = max(self.INV, key=lambda v: self.rates.get(v, 0.0))

    def step(self, seconds):
        """One bounded sample of the workload: a forward leg and an inverse leg on every thread.
        Returns single-direction NTTs per second over the pair."""
        f = self.time_variant(self.fwd_name, seconds)
        i = self.time_variant(self.inv_name, seconds)
        return 2.0 / (1.0 / f + 1.0 / i)

    def describe(self, value, seconds):
        r = self.rates
        return {
            "value": value, "unit": UNIT, "cores": self.threads, "kind": "reference",
            "fwd_variant": self.fwd_name, "inv_variant": self.inv_name,
            "fwd_ntt_per_s": r[self.fwd_name], "inv_ntt_per_s": r[self.inv_name],
            "exact_ifma_fwd_ntt_per_s": r.get("fwd_ntt_r2_16_avx512_ifma"),
            "scalar_oracle_fwd_ntt_per_s": r.get("fwd_ntt_ref_harvey"),
            "sample": "one polynomial per pthread on %d threads, %d forward + %d inverse calls per thread per "
                      "sample (about %.1f s per leg), N=2^%d, q=%#x, reference sources compiled -O3 (oracle/_ref)"
                      % (self.threads, self.calls[self.fwd_name], self.calls[self.inv_name], seconds, self.logn,
                         Q49),
            "note": "radix-4 AVX512-IFMA is the reference's fastest forward but is not exact at 49-bit q "
                    "(SURVEY.md Appendix F); fwd_ntt_r2_16_avx512_ifma is the exact IFMA figure",
        }


def cpu_reference_rates(logn, seconds):
    """cpu_baseline object for the B200 arm's JSON line (rank 0, N=1): one bounded sample."""
    cpu = CpuReference(logn)
    if not cpu.available:
        return None
    cpu.survey(seconds)
    return cpu.describe(cpu.step(seconds), seconds)


def run_reference_arm(args):
    """--impl reference: the reference's CPU path on this box's host cores, same metric and config."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    cpu = CpuReference(args.logn)
    if not cpu.available:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (needs /root/reference)"}))
        return 0
    # keep the whole run within a few minutes whatever K and W are
    seconds = max(0.1, min(1.5, 90.0 / (2.0 * (args.steps + args.warmup + 4))))
    cpu.survey(seconds)
    for _ in range(args.warmup):
        cpu.step(seconds)
    t0 = time.time()
    samples = [cpu.step(seconds) for _ in range(args.steps)]
    wall = time.time() - t0
    value = sum(samples) / len(samples)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall * 1e3 / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(args, {"device": "host CPU, %d threads" % cpu.threads}),
        "cpu_baseline": cpu.describe(value, seconds),
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---- B200 arm ----------------------------------------------------------------------------------------------

def load_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def load_compute_bound(fwd_ntt_per_s):
    """The arithmetic-pipe roofline of the forward kernel: measured DFMA issue rate (tools/ubench_pipes.cu) divided
    by the FP64 instructions one transform executes; north_star's roofline is the slower of HBM and this.
    (The integer formulation's bound, from tools/ubench_bfly.cu, is reported beside it.)"""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            t = json.load(fh)
        peak = float(t["fp64_ops_per_s"]) / float(t["fp64_ops_per_fwd_ntt_logn14"])
        return {"bound": "FP64 pipe (DFMA/DADD/DMUL issue rate)", "peak_ntt_per_s": peak,
                "achieved_ntt_per_s": fwd_ntt_per_s, "frac": fwd_ntt_per_s / peak, "source": t.get("fp64_source"),
                "integer_path_peak_ntt_per_s": float(t["int_pipe_bfly_per_s"]) / float(t["bfly_per_fwd_ntt_logn14"])}
    except Exception:
        return None


def load_traffic(logn):
    """dram bytes per launch of the forward chunk kernel from the committed ncu capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            t = json.load(fh)
        return t.get("fwd_logn%d_dram_bytes_per_ntt" % logn)
    except Exception:
        return None


def run_b200_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    ntt = importlib.import_module(PKG)
    sharding = importlib.import_module(PKG + ".sharding")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    N, batch = 1 << args.logn, args.batch
    plan = ntt.Plan.from_psi(N, Q49, psi_for(args.logn, ntt), device=local)

    # synthetic input: splitmix64 % q generated on the host by the product's own helper-free numpy code
    rng = np.random.default_rng(1 + rank)
    host = rng.integers(0, Q49, size=(batch, N), dtype=np.uint64)
    pinned = torch.from_numpy(host.view(np.int64)).pin_memory()
    dev = pinned.cuda(non_blocking=False)
    stream = torch.cuda.current_stream()

    def step():
        plan.fwd(dev, batch, stream)
        plan.inv(dev, batch, stream)

    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize()
    # correctness guard inside the bench: the round trip must reproduce the input exactly
    if not torch.equal(dev, pinned.cuda()):
        raise SystemExit("bench.py: forward+inverse round trip is not the identity")

    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    sampler = ClockSampler(local)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record(stream)
    for k in range(args.steps):
        ev[k][0].record(stream)
        plan.fwd(dev, batch, stream)
        ev[k][1].record(stream)
        plan.inv(dev, batch, stream)
        ev[k][2].record(stream)
    t_end.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    total_ms = sharding.reduce_max(t_start.elapsed_time(t_end), dist if world > 1 else None)
    fwd_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / args.steps
    inv_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / args.steps

    ms_per_step = total_ms / args.steps
    value = 2.0 * batch * world / (ms_per_step * 1e-3)

    # end to end through the host-buffer C-ABI: pinned host memory, H2D + kernels + D2H inside the timing
    e2e_steps = max(2, min(args.steps, 5))
    plan.fwd_host(pinned, batch)
    plan.inv_host(pinned, batch)  # warm-up (creates the staging pipeline)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        plan.fwd_host(pinned, batch)
        plan.inv_host(pinned, batch)
    e2e_s = sharding.reduce_max(time.perf_counter() - t0, dist if world > 1 else None) / e2e_steps
    if not np.array_equal(pinned.numpy().view(np.uint64), host):
        raise SystemExit("bench.py: host-path round trip is not the identity")
    bytes_one_way = batch * N * 8
    e2e = {"value": 2.0 * batch * world / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 2 * bytes_one_way,
           "d2h_bytes_per_step": 2 * bytes_one_way, "ms_per_step": e2e_s * 1e3,
           "api": "ntt_b200_fwd_batch_host + ntt_b200_inv_batch_host, pinned host buffers"}

    peak, peak_src = load_peak()
    alg_bytes = 2.0 * N * 8 * batch  # per launch of the forward kernel: read + write every coefficient once
    achieved = alg_bytes / (fwd_ms * 1e-3) / 1e9
    traffic = load_traffic(args.logn)
    roofline = {
        "bound": "hbm", "kernel": "k_ring_fp<%d,fwd> (one launch = %d transforms)" % (args.logn, batch),
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": None if traffic is None else traffic * batch, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": fwd_ms,
        "compute_bound": load_compute_bound(batch / (fwd_ms * 1e-3)),
        "inverse": {"kernel_ms": inv_ms, "achieved": alg_bytes / (inv_ms * 1e-3) / 1e9,
                    "frac": alg_bytes / (inv_ms * 1e-3) / 1e9 / peak},
        "fwd_ntt_per_s_per_gpu": batch / (fwd_ms * 1e-3), "inv_ntt_per_s_per_gpu": batch / (inv_ms * 1e-3),
    }

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference_rates(args.logn, args.cpu_seconds)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, {"arithmetic": "u64 coefficients; butterflies are exact integer arithmetic "
                                                           "carried in FP64 (q < 2^50), results bit-identical to the "
                                                           "reference's 64-bit integer code"}),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": 2 * args.steps * world, "clocks": clocks, "impl": "b200",
        }
        print(json.dumps(line))
    plan.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())
