#!/usr/bin/env python
"""bench.py -- throughput of the negacyclic NTT hot path on B200 (one process per GPU).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json configs[1]): batched forward+inverse NTT, N = 2^14, 49-bit q = 0x1fffffc800001,
4096 polynomials per GPU, synthetic uniform coefficients (splitmix64 % q).  One step = forward transform
of the whole batch followed by the inverse transform of the whole batch (2 * batch single-direction NTTs
per GPU).  The batch is 512 MiB per GPU, four times the 126 MB L2, so every step streams from HBM.

Printed JSON (rank 0, one line):
  metric/value   whole-job single-direction NTTs per second over all GPUs, CUDA events on the launch stream,
                 max over ranks (`sustained`: the same loop run for >= 2 s with its own clock record);
  roofline       forward chunk kernel against the measured HBM copy bandwidth (algorithmic bytes 2*N*8 per
                 transform), the inverse beside it, and `issue_bound`: the SM issue-slot roofline that actually
                 limits the FP64 formulation;
  cpu_baseline   the reference's own CPU code (oracle/_ref) on this box's host cores (N=1 only);
  e2e            same metric through the host-buffer C-ABI call ntt_b200_fwd_mul_inv_batch_host (every chunk
                 crosses PCIe once per direction for a forward AND an inverse transform), pinned host memory,
                 copies inside the timing; `separate_calls` = forward and inverse as two host calls (the round-1
                 figure); `copy_bound` = a bare pinned H2D+D2H of the same bytes on the same box;
  parity         >= 8 random polynomials of the timed batch compared with the oracle, outside the timed region;
  other_configs  BASELINE configs 3 (RNS N=2^16 x 48 limbs, sharded by limb), 4 (negacyclic multiply N=2^13) and,
                 when WORLD_SIZE > 1, 5 (one N=2^22 transform over all GPUs, exchange over NVLink).

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
PSI = {13: 94912374482, 14: 20456969886, 16: 3471868370, 22: 142690821}  # smallest primitive 2N-th roots
BATCH_PER_GPU = 4096
METRIC = "fwd+inv NTTs/s (N=2^14, 49-bit q, batched)"
UNIT = "NTT/s"
PARITY_ROWS = 8


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="polynomials per GPU")
    ap.add_argument("--logn", type=int, default=LOGN)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip other_configs, sustained and copy-bound legs")
    ap.add_argument("--cpu-seconds", type=float, default=3.0, help="target seconds per CPU baseline leg")
    ap.add_argument("--sustain-seconds", type=float, default=2.0)
    return ap.parse_args()


def workload_config(args):
    """Identical in both arms (the driver compares the dicts): the workload only, nothing about who runs it."""
    return {
        "workload": "batched forward+inverse negacyclic NTT, N=2^%d, 49-bit q=%#x, batch %d per GPU "
                    "(BASELINE configs[1])" % (args.logn, Q49, args.batch),
        "N": 1 << args.logn, "q": Q49, "batch_per_gpu": args.batch,
        "input": "splitmix64(seed = 1 + rank) %% q, uniform in [0,q)",
        "step": "forward then inverse transform of the batch; value counts single-direction transforms",
        "cache": "inputs larger than L2 (batch is %d MiB per GPU)" % ((args.batch << args.logn) * 8 >> 20),
        "parallelism": "polynomials sharded across GPUs, no data-path collective",
    }


def psi_for(logn, ntt=None):
    if logn in PSI:
        return PSI[logn]
    return ntt.min_primitive_root(1 << logn, Q49)


def splitmix64_mod(n, q, seed):
    """a[i] = splitmix64 stream (state seed) reduced mod q -- vectorised, bit-identical
    to oracle_fill_uniform / the generator behind the golden hashes."""
    import numpy as np
    out = np.empty(n, dtype=np.uint64)
    step = 1 << 22
    with np.errstate(over="ignore"):
        for lo in range(0, n, step):
            k = np.arange(lo + 1, min(n, lo + step) + 1, dtype=np.uint64)
            z = np.uint64(seed) + k * np.uint64(0x9E3779B97F4A7C15)
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            out[lo:lo + len(k)] = (z ^ (z >> np.uint64(31))) % np.uint64(q)
    return out


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


def bind_near_gpu(index):
    """Run this process (and so allocate its pinned buffers) on the CPUs NVML reports as local to the GPU."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index))
        return True
    except Exception:
        return False


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
        "config": workload_config(args),
        "arm": "reference CPU code on the host, %d threads" % cpu.threads,
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


def load_profile_numbers():
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            return json.load(fh)
    except Exception:
        return {}


def issue_bound(prof, fwd_ntt_per_s, sm_mhz):
    """The roofline that actually binds the FP64 formulation: the SM issue slots.  An FP64 instruction holds a
    scheduler's dispatch port for 2 cycles and every other instruction for 1 (tools/ubench_rf.cu,
    profiles/r02_ubench_rf.txt), so one transform costs (2*F + O) issue cycles per warp, F and O counted from the
    SASS of the shipped kernel (profiles/traffic.json)."""
    try:
        f, o = float(prof["fwd14_fp64_instr_per_thread"]), float(prof["fwd14_other_instr_per_thread"])
        warps, scheds, sms = 16.0, 4.0, 148.0
        cycles = (2.0 * f + o) * warps / scheds                 # per polynomial per SM
        peak = sms * (sm_mhz or 1965) * 1e6 / cycles
        return {"bound": "SM issue slots (2 cycles per FP64 instruction, 1 per other instruction)",
                "fp64_instr_per_thread": f, "other_instr_per_thread": o, "peak_ntt_per_s": peak,
                "achieved_ntt_per_s": fwd_ntt_per_s, "frac": fwd_ntt_per_s / peak,
                "source": prof.get("issue_source")}
    except Exception:
        return None


def event_pair(torch):
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed_ms(torch, fn, steps, warm=3, stream=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = event_pair(torch)
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def oracle_tables(orc, logn, q, psi):
    N = 1 << logn
    w, wc = orc.tables(N, q, psi)
    psi_inv = orc.invmod(psi, q)
    wi, wic = orc.tables(N, q, psi_inv)
    return dict(w=w, wc=wc, wi=wi, wic=wic, n_inv=orc.invmod(N, q))


def run_config3(ntt, torch, dist, rank, world, local, peak):
    """BASELINE config 3: CKKS-style RNS batch, N = 2^16, 48 limbs (largest 49-bit primes = 1 mod 2^17), 32
    polynomials per limb, limbs sharded contiguously over the ranks; one limb spot-checked against the oracle."""
    import numpy as np
    from oracle.pyoracle import Oracle
    m, limbs, per = 16, 48, 32
    N = 1 << m
    qs, q = [], (1 << 49) + 1
    q -= (q - 1) % (2 * N)
    while len(qs) < limbs:
        q -= 2 * N
        if ntt.is_prime(q) and q <= (1 << 49) - 1024:
            qs.append(q)
    lb, le = limbs * rank // world, limbs * (rank + 1) // world
    mine = qs[lb:le]
    psis = [ntt.min_primitive_root(N, ql) for ql in mine]
    plans = [ntt.Plan.from_psi(N, ql, ps, device=local) for ql, ps in zip(mine, psis)]
    a = np.stack([splitmix64_mod(per * N, ql, 2000 + lb + i).reshape(per, N) for i, ql in enumerate(mine)])
    d = torch.from_numpy(a.view(np.int64)).cuda()
    ntt.fwd_rns(plans, d, per)
    f = d.cpu().numpy().view(np.uint64)
    orc = Oracle()
    t = oracle_tables(orc, m, mine[0], psis[0])
    ok = bool(np.array_equal(f[0, :2], orc.fwd_batch(a[0, :2], mine[0], t["w"], t["wc"])))
    ntt.inv_rns(plans, d, per)
    ok = ok and bool(np.array_equal(d.cpu().numpy().view(np.uint64), a))
    if world > 1:
        dist.barrier()
    ms_f = timed_ms(torch, lambda: ntt.fwd_rns(plans, d, per), 10)
    ms_i = timed_ms(torch, lambda: ntt.inv_rns(plans, d, per), 10)
    if world > 1:
        tt = torch.tensor([ms_f, ms_i, 0.0 if ok else 1.0], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_f, ms_i, bad = tt.tolist()
        ok = bad == 0.0
    for p in plans:
        p.close()
    n = limbs * per
    return {"config": "RNS N=2^16 x %d limbs x %d polynomials, 49-bit primes, limbs sharded over %d GPU(s)"
                      % (limbs, per, world),
            "fwd_ms": ms_f, "inv_ms": ms_i, "fwd_ntt_per_s": n / ms_f * 1e3, "inv_ntt_per_s": n / ms_i * 1e3,
            "fwd_frac_of_hbm_single_pass_per_gpu": n * 2 * N * 8 / (ms_f * 1e-3) / (peak * 1e9) / world,
            "inv_frac_of_hbm_single_pass_per_gpu": n * 2 * N * 8 / (ms_i * 1e-3) / (peak * 1e9) / world,
            "parity_vs_oracle": ok}


def run_config4(ntt, torch, peak):
    """BASELINE config 4: negacyclic polynomial multiply, N = 2^13, batch 16384; rows checked against the oracle
    pipeline (forward x2, pointwise product, inverse)."""
    import numpy as np
    from oracle.pyoracle import Oracle
    m, batch, q, psi = 13, 16384, Q49, PSI[13]
    N = 1 << m
    plan = ntt.Plan.from_psi(N, q, psi)
    a = splitmix64_mod(batch * N, q, 3).reshape(batch, N)
    b = splitmix64_mod(batch * N, q, 33).reshape(batch, N)
    da, db = torch.from_numpy(a.view(np.int64)).cuda(), torch.from_numpy(b.view(np.int64)).cuda()
    dc = torch.empty_like(da)
    plan.negacyclic_mul(dc, da, db, batch)
    c = dc.cpu().numpy().view(np.uint64)
    orc = Oracle()
    t = oracle_tables(orc, m, q, psi)
    rows = [0, 1, 4097, 9999, batch - 1]
    fa, fb = orc.fwd_batch(a[rows], q, t["w"], t["wc"]), orc.fwd_batch(b[rows], q, t["w"], t["wc"])
    want = orc.inv_batch(orc.pointwise_mul(fa, fb, q).reshape(len(rows), N), q, t["n_inv"], t["wi"], t["wic"])
    ok = bool(np.array_equal(c[rows], want))
    da.copy_(torch.from_numpy(a.view(np.int64)))
    db.copy_(torch.from_numpy(b.view(np.int64)))
    ms = timed_ms(torch, lambda: plan.negacyclic_mul(da, da, db, batch), 10)   # in place: db is work space
    plan.close()
    return {"config": "negacyclic polynomial multiply N=2^13, batch %d" % batch, "ms": ms,
            "products_per_s": batch / ms * 1e3, "frac_of_hbm_3N8": batch * 3 * N * 8 / (ms * 1e-3) / (peak * 1e9),
            "parity_vs_oracle": ok}


def run_config5(ntt, torch, dist, rank, world, local):
    """BASELINE config 5: ONE forward+inverse NTT of size N = 2^22 spread over all ranks -- the only path with a real
    exchange.  Two variants: NCCL all-to-all between the local transforms and the tail stages, and the exchange
    fused into the tail kernels over NVLink peer memory (CUDA IPC, GPU-side flag barrier).  Rank blocks of the
    forward transform are gathered and compared with the oracle; timings are device events, max over ranks."""
    import numpy as np
    fs = importlib.import_module(PKG + ".fourstep")
    from oracle.pyoracle import Oracle
    m = 22
    N, q, psi = 1 << m, Q49, PSI[22]
    a = splitmix64_mod(N, q, 4)
    steps = 20
    out = {"config": "one N=2^22 forward+inverse transform over %d GPUs" % world, "N": N,
           "bytes_exchanged_per_gpu_per_direction": (N // world) * 8 * (world - 1) // world}

    def reduce_ms(ms):
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # single-GPU time of the same transform (every rank measures its own GPU; max taken)
    single = ntt.Plan.from_psi(N, q, psi, device=local)
    ds = torch.from_numpy(a.view(np.int64)).cuda()
    out["one_gpu_ms_per_pair"] = reduce_ms(timed_ms(torch, lambda: (single.fwd(ds, 1), single.inv(ds, 1)), steps))
    single.close()
    del ds

    # NCCL all-to-all variant
    plan = fs.DistributedNtt(N, q, psi, rank, world, device=local)
    sl0 = torch.from_numpy(np.ascontiguousarray(a[rank::world]).view(np.int64)).cuda()
    blk = plan.forward(sl0.clone(), dist)
    gathered = [torch.empty_like(blk) for _ in range(world)] if rank == 0 else None
    dist.gather(blk, gathered, dst=0)
    ok_fwd = None
    if rank == 0:
        orc = Oracle()
        t = oracle_tables(orc, m, q, psi)
        want = orc.fwd(a, q, t["w"], t["wc"])
        ok_fwd = bool(np.array_equal(torch.cat(gathered).cpu().numpy().view(np.uint64), want))
    back = plan.inverse(blk, dist)
    ok_rt = bool(torch.equal(back, sl0))

    def step_nccl(state=[sl0.clone()]):
        state[0] = plan.inverse(plan.forward(state[0], dist), dist)
    dist.barrier()
    out["nccl_ms_per_pair"] = reduce_ms(timed_ms(torch, step_nccl, steps))
    plan.close()

    # exchange fused into the tail kernels
    fused = fs.FusedDistributedNtt(N, q, psi, rank, world, local, dist)
    fused.px.load_slice(a[rank::world])
    block = torch.empty(N // world, dtype=torch.int64, device="cuda")
    fused.forward(block)
    torch.cuda.synchronize()
    gathered = [torch.empty_like(block) for _ in range(world)] if rank == 0 else None
    dist.gather(block, gathered, dst=0)
    ok_fused = None
    if rank == 0:
        ok_fused = bool(np.array_equal(torch.cat(gathered).cpu().numpy().view(np.uint64), want))
    fused.inverse(block)
    torch.cuda.synchronize()
    ok_rt = ok_rt and bool(np.array_equal(fused.px.read_slice(), a[rank::world])) and not fused.px.timed_out()

    def step_fused():
        fused.forward(block)
        fused.inverse(block)
    dist.barrier()
    out["peer_fused_ms_per_pair"] = reduce_ms(timed_ms(torch, step_fused, steps))
    graph = fused.capture_pair(block)
    dist.barrier()
    out["peer_fused_graph_ms_per_pair"] = reduce_ms(timed_ms(torch, graph.replay, steps))
    ok_rt = ok_rt and bool(np.array_equal(fused.px.read_slice(), a[rank::world])) and not fused.px.timed_out()
    del graph
    fused.close()

    # the same exchange with a BATCH of transforms per launch / barrier: the single transform is latency-bound
    # (32 MB over 8 GPUs), a batch amortises the launches and the barrier and fills the SMs
    # 9 transforms = 288 chunks of 2^14 per rank at 8 ranks: two full waves of the 148 persistent CTAs (8 would leave a
    # quarter of the second wave empty)
    B = 9
    ab = np.stack([a] + [splitmix64_mod(N, q, 40 + i) for i in range(1, B)])
    fb = fs.FusedDistributedNtt(N, q, psi, rank, world, local, dist, batch=B)
    fb.px.load_slice(np.ascontiguousarray(ab[:, rank::world]).reshape(-1))
    blocks = torch.empty(B * (N // world), dtype=torch.int64, device="cuda")
    fb.forward(blocks)
    torch.cuda.synchronize()
    first = blocks[:N // world].clone()                      # polynomial 0 of the batch is `a`
    gathered = [torch.empty_like(first) for _ in range(world)] if rank == 0 else None
    dist.gather(first, gathered, dst=0)
    ok_batched = None
    if rank == 0:
        ok_batched = bool(np.array_equal(torch.cat(gathered).cpu().numpy().view(np.uint64), want))
    fb.inverse(blocks)
    torch.cuda.synchronize()
    ok_rt = ok_rt and bool(np.array_equal(fb.px.read_slice(), np.ascontiguousarray(ab[:, rank::world]).reshape(-1)))

    def step_batched():
        fb.forward(blocks)
        fb.inverse(blocks)
    dist.barrier()
    out["peer_fused_batch%d_ms_per_pair_per_polynomial" % B] = reduce_ms(timed_ms(torch, step_batched, steps)) / B
    ok_rt = ok_rt and not fb.px.timed_out()
    fb.close()
    single_b = ntt.Plan.from_psi(N, q, psi, device=local)
    dsb = torch.from_numpy(ab.view(np.int64)).cuda()
    out["one_gpu_batch%d_ms_per_pair_per_polynomial" % B] = reduce_ms(
        timed_ms(torch, lambda: (single_b.fwd(dsb, B), single_b.inv(dsb, B)), steps)) / B
    single_b.close()
    out["batched_speedup_over_one_gpu"] = (out["one_gpu_batch%d_ms_per_pair_per_polynomial" % B]
                                           / out["peer_fused_batch%d_ms_per_pair_per_polynomial" % B])
    out["forward_blocks_equal_oracle_peer_fused_batched"] = ok_batched
    bad = torch.tensor([0.0 if ok_rt else 1.0], device="cuda")
    dist.all_reduce(bad, op=dist.ReduceOp.MAX)
    best = min(out["peer_fused_ms_per_pair"], out["peer_fused_graph_ms_per_pair"])
    out.update({
        "forward_blocks_equal_oracle_nccl": ok_fwd, "forward_blocks_equal_oracle_peer_fused": ok_fused,
        "round_trip_identity_all_ranks": bad.item() == 0.0,
        "speedup_over_one_gpu": out["one_gpu_ms_per_pair"] / best,
        # two exchanges per pair (forward gather, inverse scatter), each moving bytes_exchanged per GPU per direction
        "nvlink_GBps_per_gpu_if_exchange_took_the_whole_pair": 2 * out["bytes_exchanged_per_gpu_per_direction"]
                                                               / (best * 1e-3) / 1e9,
    })
    return out


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
    all_cpus = os.sched_getaffinity(0)
    near = bind_near_gpu(local)                      # pinned buffers on the GPU's own NUMA node where NVML knows it
    host_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        host_group = dist.new_group(backend="gloo")  # host-side barriers that put no kernel on the GPUs
    D = dist if world > 1 else None

    N, batch = 1 << args.logn, args.batch
    psi = psi_for(args.logn, ntt)
    plan = ntt.Plan.from_psi(N, Q49, psi, device=local)
    fwd_kernels, fwd_launches = plan.describe(False)
    inv_kernels, inv_launches = plan.describe(True)

    host = splitmix64_mod(batch * N, Q49, 1 + rank).reshape(batch, N)
    pinned = torch.from_numpy(host.view(np.int64)).pin_memory()
    dev = pinned.cuda(non_blocking=False)
    stream = torch.cuda.current_stream()

    def step():
        plan.fwd(dev, batch, stream)
        plan.inv(dev, batch, stream)

    # ---- parity, outside the timed region: PARITY_ROWS random polynomials against the oracle, both directions ----
    from oracle.pyoracle import Oracle
    orc = Oracle()
    tb = oracle_tables(orc, args.logn, Q49, psi)
    rows = sorted(set(np.random.default_rng(99 + rank).integers(0, batch, size=PARITY_ROWS).tolist()) | {0, batch - 1})
    plan.fwd(dev, batch, stream)
    torch.cuda.synchronize()
    fwd_rows = dev[rows].cpu().numpy().view(np.uint64)
    fwd_ok = bool(np.array_equal(fwd_rows, orc.fwd_batch(host[rows], Q49, tb["w"], tb["wc"])))
    plan.inv(dev, batch, stream)
    torch.cuda.synchronize()
    inv_ok = bool(np.array_equal(dev[rows].cpu().numpy().view(np.uint64),
                                 orc.inv_batch(fwd_rows, Q49, tb["n_inv"], tb["wi"], tb["wic"])))
    rt_ok = bool(torch.equal(dev, pinned.cuda()))
    if not (fwd_ok and inv_ok and rt_ok):
        raise SystemExit("bench.py: parity check failed (forward %s, inverse %s, round trip %s)" % (fwd_ok, inv_ok, rt_ok))
    parity = {"rows_vs_oracle": len(rows), "forward": fwd_ok, "inverse": inv_ok, "round_trip_whole_batch": rt_ok}

    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize()

    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    sampler = ClockSampler(local)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    t_start, t_end = event_pair(torch)
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
    total_ms = sharding.reduce_max(t_start.elapsed_time(t_end), D)
    fwd_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / args.steps
    inv_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / args.steps
    ms_per_step = total_ms / args.steps
    value = 2.0 * batch * world / (ms_per_step * 1e-3)
    if not torch.equal(dev, pinned.cuda()):
        raise SystemExit("bench.py: the batch changed over the timed steps (round trip is not the identity)")

    # ---- sustained leg: the same step back to back for >= sustain-seconds, its own clock record ------------------
    sustained = None
    if not args.no_extras and args.sustain_seconds > 0:
        n_sus = max(args.steps, int(args.sustain_seconds * 1e3 / ms_per_step) + 1)
        s2 = ClockSampler(local)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        s2.start()
        a0, a1 = event_pair(torch)
        a0.record(stream)
        for _ in range(n_sus):
            step()
        a1.record(stream)
        torch.cuda.synchronize()
        c2 = s2.stop()
        sus_ms = sharding.reduce_max(a0.elapsed_time(a1), D) / n_sus
        sustained = {"value": 2.0 * batch * world / (sus_ms * 1e-3), "unit": UNIT, "steps": n_sus,
                     "seconds": sus_ms * n_sus * 1e-3, "ms_per_step": sus_ms, "clocks": c2}

    # ---- end to end through the host-buffer C-ABI ------------------------------------------------------------
    bytes_one_way = batch * N * 8
    mult = splitmix64_mod(N, Q49, 777)                                  # NTT-domain multiplier, resident on the GPU
    d_mult = torch.from_numpy(mult.view(np.int64)).cuda()
    e2e_steps = max(2, min(args.steps, 5))
    track = host[rows].copy()                                           # the oracle follows these rows step by step

    def oracle_fmi(x):
        f = orc.fwd_batch(x, Q49, tb["w"], tb["wc"])
        p = orc.pointwise_mul(f, np.ascontiguousarray(np.broadcast_to(mult, f.shape)), Q49).reshape(f.shape)
        return orc.inv_batch(p, Q49, tb["n_inv"], tb["wi"], tb["wic"])

    plan.fwd_mul_inv_host(pinned, d_mult, batch)                        # warm-up (creates the staging pipeline)
    track = oracle_fmi(track)
    if world > 1:
        dist.barrier(group=host_group)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        plan.fwd_mul_inv_host(pinned, d_mult, batch)
    e2e_s = sharding.reduce_max(time.perf_counter() - t0, D) / e2e_steps
    for _ in range(e2e_steps):
        track = oracle_fmi(track)
    e2e_ok = bool(np.array_equal(pinned.numpy().view(np.uint64)[rows], track))
    if not e2e_ok:
        raise SystemExit("bench.py: host-path result differs from the oracle pipeline")
    parity["e2e_rows_vs_oracle_after_%d_calls" % (e2e_steps + 1)] = e2e_ok
    # the round-1 form: forward and inverse as two host calls (each crosses PCIe both ways)
    pinned.copy_(torch.from_numpy(host.view(np.int64)))
    plan.fwd_host(pinned, batch)
    plan.inv_host(pinned, batch)
    if world > 1:
        dist.barrier(group=host_group)
    t0 = time.perf_counter()
    for _ in range(2):
        plan.fwd_host(pinned, batch)
        plan.inv_host(pinned, batch)
    sep_s = sharding.reduce_max(time.perf_counter() - t0, D) / 2
    if not np.array_equal(pinned.numpy().view(np.uint64), host):
        raise SystemExit("bench.py: host-path round trip is not the identity")
    e2e = {"value": 2.0 * batch * world / e2e_s, "unit": UNIT, "h2d_bytes_per_step": bytes_one_way,
           "d2h_bytes_per_step": bytes_one_way, "ms_per_step": e2e_s * 1e3,
           "api": "ntt_b200_fwd_mul_inv_batch_host (forward, NTT-domain product with a resident polynomial, inverse; "
                  "one H2D and one D2H per polynomial), pinned host buffers%s" % (", CPU affinity set near the GPU" if near else ""),
           "separate_calls": {"value": 2.0 * batch * world / sep_s, "ms_per_step": sep_s * 1e3,
                              "h2d_bytes_per_step": 2 * bytes_one_way, "d2h_bytes_per_step": 2 * bytes_one_way,
                              "api": "ntt_b200_fwd_batch_host + ntt_b200_inv_batch_host"}}
    # what the box can copy: bare pinned H2D + D2H of the same bytes, both directions at once, all ranks together
    if not args.no_extras:
        other = torch.empty_like(pinned).pin_memory()
        dev2 = torch.empty_like(dev)
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

        def copy_both():
            with torch.cuda.stream(s_in):
                dev.copy_(pinned, non_blocking=True)
            with torch.cuda.stream(s_out):
                other.copy_(dev2, non_blocking=True)
        copy_both()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=host_group)
        t0 = time.perf_counter()
        for _ in range(3):
            copy_both()
        torch.cuda.synchronize()
        cp_s = sharding.reduce_max(time.perf_counter() - t0, D) / 3
        e2e["copy_bound"] = {"ms_per_step": cp_s * 1e3, "GBps_per_direction_per_gpu": bytes_one_way / cp_s / 1e9,
                             "value_if_copies_were_all": 2.0 * batch * world / cp_s,
                             "what": "pinned cudaMemcpyAsync H2D and D2H of one step's bytes, concurrently, all ranks at once"}
        e2e["frac_of_copy_bound"] = cp_s / e2e_s
        del other, dev2

    peak, peak_src = load_peak()
    prof = load_profile_numbers()
    alg_bytes = 2.0 * N * 8 * batch  # per launch of the forward kernel: read + write every coefficient once
    achieved = alg_bytes / (fwd_ms * 1e-3) / 1e9
    traffic = prof.get("fwd_logn%d_batch%d_dram_bytes_per_launch" % (args.logn, batch))
    roofline = {
        "bound": "hbm", "kernel": "%s (one launch = %d transforms)" % (fwd_kernels, batch),
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic, "traffic_source": prof.get("traffic_source") if traffic else None,
        "peak_source": peak_src,
        "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": fwd_ms,
        "issue_bound": issue_bound(prof, batch / (fwd_ms * 1e-3), clocks.get("sm_mhz")),
        "inverse": {"kernel": inv_kernels, "kernel_ms": inv_ms, "achieved": alg_bytes / (inv_ms * 1e-3) / 1e9,
                    "frac": alg_bytes / (inv_ms * 1e-3) / 1e9 / peak},
        "fwd_ntt_per_s_per_gpu": batch / (fwd_ms * 1e-3), "inv_ntt_per_s_per_gpu": batch / (inv_ms * 1e-3),
    }

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        os.sched_setaffinity(0, all_cpus)            # the CPU baseline gets every host core again
        cpu = cpu_reference_rates(args.logn, args.cpu_seconds)

    plan.close()
    del dev, pinned
    others = {}
    if not args.no_extras:
        torch.cuda.empty_cache()
        try:
            others["config3_rns"] = run_config3(ntt, torch, dist, rank, world, local, peak)
        except Exception as e:  # the headline line must survive a failure in a side config
            others["config3_rns"] = {"error": repr(e)}
        if rank == 0:
            try:
                others["config4_polymul"] = run_config4(ntt, torch, peak)
            except Exception as e:
                others["config4_polymul"] = {"error": repr(e)}
        if world > 1:
            dist.barrier()
            try:
                others["config5_large_n"] = run_config5(ntt, torch, dist, rank, world, local)
            except Exception as e:
                others["config5_large_n"] = {"error": repr(e)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args),
            "arm": "sm_100a kernels; u64 coefficients, butterflies are exact integer arithmetic carried in FP64 "
                   "(q < 2^50), results bit-identical to the reference's 64-bit integer code",
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "sustained": sustained, "parity": parity,
            "gpu_launches": (fwd_launches + inv_launches) * args.steps * world,
            "kernels": {"forward": fwd_kernels, "inverse": inv_kernels},
            "clocks": clocks, "impl": "b200", "other_configs": others,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    # stdout carries exactly ONE JSON line: libraries that print there (NCCL's version banner, torchrun notices)
    # are sent to stderr for the duration of the run and the line is written to the real stdout at the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import builtins
    import io
    out = io.TextIOWrapper(os.fdopen(real_stdout, "wb"), write_through=True)
    py_print = builtins.print

    def print_json(*a, **k):
        k.setdefault("file", out)
        py_print(*a, **k)
    global print
    print = print_json
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())
