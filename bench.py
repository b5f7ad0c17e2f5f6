#!/usr/bin/env python
"""bench.py - throughput of the COVIDCurve hot path on B200 (and of the CPU reference arm).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--vectors V]

Likelihood leg (BASELINE.json configs[4] = "cfg5", the likelihood-only microbenchmark on the configs[2] shape the metric is
quoted on): V = 2^20 random parameter vectors of the 10-age-band / 300-day / 5-knot / 3-serosurvey binomial model (d = 34),
synthetic data (covidcurve_b200/synth.py, seeds 20201/20202).  One "step" = one pass of the batched log-likelihood over all V
vectors, SPLIT EVENLY over the N ranks (strong scaling, V / N vectors per GPU; no collective).
  value  = log-likelihood evaluations per second, whole job, inputs resident in HBM;
  weak   = the same with V vectors on EVERY GPU (second figure, N > 1 only);
  e2e    = `value`'s workload through the C ABI with HOST (pinned) buffers: host->device copy of the parameter vectors and
           device->host copy of the log-likelihoods inside the timed region; `e2e.pageable` = the same from plain malloc'ed
           memory (what an R caller's REAL(theta) is).
MCMC legs (GPU-resident Metropolis-coupled MCMC; one iteration = one sweep of d single-site updates + the coupling pass;
figures are iterations/s summed over all chains x rungs; device time from CUDA events on the library's stream):
  mcmc       = BASELINE configs[3] "cfg4": 17 age bands x 400 days, 4096 chains x 64 rungs SHARDED over the N GPUs
               (4096 / N chains per GPU, RNG keyed by global chain id), followed by the run's only collective:
               an NCCL all-gather of the cold-rung sampling draws + the 3 x d all-reduce form of R-hat (`collective`);
  mcmc_cfg3  = BASELINE configs[2] "cfg3": 64 chains x 50 rungs on one GPU (replicated per GPU when N > 1).
Both check their final cold-rung states against the CPU oracle.  cpu_baseline / --impl reference time the oracle port of the
reference (likelihood and the drjacoby-equivalent MCMC loop) on the host cores: 1 thread and all threads.

Multi-GPU: launched by torchrun, one rank per GPU; timing = max over ranks of CUDA-event time between barriers.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from covidcurve_b200 import synth  # noqa: E402

METRIC = "loglike_evals_per_sec"
UNIT = "evals/s"
WORKLOAD = "cfg5: likelihood-only, 10 age bands x 300 days x 5 knots x 3 serosurveys, binomial (d=34)"


def f_alg(spec):
    """Algorithmic flops per full evaluation: T(T+1) + serorev*M(M+1) + 3*S*M + 10*T."""
    T, M, S = spec.T, spec.M, spec.S
    return T * (T + 1) + (M * (M + 1) if spec.account_serorev else 0) + 3 * S * M + 10 * T


def source_digest():
    """sha256 over the CUDA sources: ties a committed ncu capture (profiles/r2_traffic.json) to the build it measured."""
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "covidcurve_b200", "csrc")
    for f in sorted(os.listdir(csrc)):
        h.update(f.encode())
        h.update(open(os.path.join(csrc, f), "rb").read())
    return h.hexdigest()[:16]


def measured_traffic():
    """dram bytes per evaluation of the dominant kernel from the committed ncu --set full capture, valid only for the build
    whose sources it was taken on (no constant survives a kernel change silently)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
    except Exception:
        return None, "no profiles/r2_traffic.json"
    if t.get("source_digest") != source_digest():
        return None, "profiles/r2_traffic.json was captured on sources %s, this build is %s: not reported" % (t.get("source_digest"), source_digest())
    return t, "profiles/r2_traffic.json (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum of %s at %d vectors)" % (
        t.get("kernel"), t.get("vectors", 0))


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md): NVML from a thread every 5 ms
    (nvidia-smi -lms cannot deliver a sample inside a 60 ms region, least of all with eight ranks polling at once);
    nvidia-smi is the fallback when the NVML binding is missing."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nv, self.stop_flag = index, [], None, None, False
        self.sm, self.mx, self.reasons, self.power = [], None, set(), []

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return self
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.reasons.update(name for bit, name in self.BITS.items() if mask & bit)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            time.sleep(0.005)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nv is not None:
            self.stop_flag = True
            self.t.join(timeout=1)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_min_mhz": float(np.min(self.sm)) if self.sm else None,
                    "sm_max_mhz": self.mx, "reasons": sorted(self.reasons), "power_w_max": float(np.max(self.power)) if self.power else None,
                    "samples": len(self.sm), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "source": "nvidia-smi"}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------ CPU arm (oracle port)
def cpu_loglike_rate(o, theta, seconds, threads):
    """evals/s of the CPU oracle (loop-for-loop restatement of the reference, g++ -O2) on `threads` host threads."""
    probe = theta[: 32 * threads]
    t0 = time.perf_counter()
    o.loglike(probe, nthreads=threads)
    rate = len(probe) / (time.perf_counter() - t0)
    n = int(max(threads * 32, min(len(theta), rate * seconds)))
    t0 = time.perf_counter()
    out = o.loglike(theta[:n], nthreads=threads)
    dt = time.perf_counter() - t0
    return n / dt, n, dt, out


def cpu_mcmc_rate(cfg, threads, seconds):
    """chain-rung iterations/s of the oracle's drjacoby-equivalent loop (oracle/cc_oracle_mcmc.cpp) on `threads` host threads:
    one chain per thread like drjacoby's `cluster=` mode, a reduced ladder (4 rungs) and as many iterations as fit the budget."""
    from oracle.oracle import Oracle
    spec, truth = synth.make_spec(cfg)
    o = Oracle(spec)
    rungs, chains = 4, threads
    init = np.tile(synth.true_theta(spec, truth), (chains, 1))
    t0 = time.perf_counter()
    o.loglike(init[:1])
    per_eval = max(time.perf_counter() - t0, 1e-4)
    iters = int(max(3, min(200, seconds / (per_eval * spec.d * rungs))))
    t0 = time.perf_counter()
    o.mcmc(burnin=iters, samples=0, rungs=rungs, chains=chains, seed=synth.SEED_MCMC, init=init, nthreads=threads)
    dt = time.perf_counter() - t0
    done = chains * rungs * (iters - 1)          # iteration 1 is the initial state
    return {"iter_per_sec_all_chains_x_rungs": done / dt, "loglike_evals_per_sec": done * spec.d / dt, "cores": threads,
            "sample": "%s: %d chains x %d rungs x %d iterations (d=%d), %.1f s" % (cfg, chains, rungs, iters - 1, spec.d, dt)}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port: the reference itself needs
    R + Rcpp + libRmath, absent from this image) with all host threads, on the same config / metric."""
    if rank != 0:
        return
    from oracle.oracle import Oracle
    spec, _ = synth.make_spec("cfg3")
    cores = host_threads()
    per_step = args.ref_vectors
    theta = synth.random_theta(spec, per_step, seed=synth.SEED_THETA, bad_frac=0.01)
    o = Oracle(spec)
    for _ in range(args.warmup):
        o.loglike(theta[: max(cores * 16, per_step // 8)], nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.loglike(theta, nthreads=cores)
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    one, n1, dt1, _ = cpu_loglike_rate(o, theta, 4.0, 1)
    sample = "%d random cfg5 vectors per step (seed %d), %d threads, g++ -O2 oracle port" % (per_step, synth.SEED_THETA, cores)
    mc = None
    if not args.no_mcmc:
        mc = {"cfg3_all_cores": cpu_mcmc_rate("cfg3", cores, 6.0), "cfg3_1_core": cpu_mcmc_rate("cfg3", 1, 5.0),
              "cfg4_all_cores": cpu_mcmc_rate("cfg4", cores, 8.0)}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "vectors_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "value_1_core": one, "sample_1_core": "%d vectors, %.1f s, 1 thread" % (n1, dt1)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "mcmc_cpu": mc,
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--vectors", type=int, default=1 << 20, help="parameter vectors per step (whole job; split over the ranks)")
    ap.add_argument("--ref-vectors", type=int, default=8192, help="vectors per step of the CPU reference arm")
    ap.add_argument("--mcmc-iters", type=int, default=-1, help="timed iterations of the cfg4 MCMC leg (-1: 25 per GPU, 12..200)")
    ap.add_argument("--mcmc-settle", type=int, default=40, help="untimed iterations before the timed ones (bandwidth adaptation)")
    ap.add_argument("--mcmc-chains", type=int, default=0, help="total chains of the cfg4 leg (0: BASELINE's 4096)")
    ap.add_argument("--cfg3-seconds", type=float, default=2.5, help="device time the cfg3 MCMC leg aims for")
    ap.add_argument("--no-mcmc", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work budget of the cpu_baseline likelihood leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-shapes", action="store_true", help="skip the cfg2 / cfg4 likelihood figures")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from covidcurve_b200 import _lib
    from covidcurve_b200 import dist as ccdist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = None
    if world > 1:
        # one process per GPU: keep this rank (and the pinned buffers it first-touches) on the CPUs next to its GPU, so that the
        # ranks' host->device copies do not cross sockets or share one memory controller
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
            numa = "cpus %d..%d" % (min(os.sched_getaffinity(0)), max(os.sched_getaffinity(0)))
        except Exception as e:  # affinity is an optimisation, never a requirement
            numa = "unset (%s)" % type(e).__name__
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_ranks(x, op="max"):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op={"max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN, "sum": dist.ReduceOp.SUM}[op])
        return float(t.item())

    spec, truth = synth.make_spec("cfg3")
    d, V = spec.d, args.vectors
    model = _lib.Model(spec, device=local_rank)
    # a real (non-default) torch stream: the library launches on the stream handle it is given, and torch's CUDA events
    # only see the stream they are recorded on
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    L2_BYTES = 126e6
    flush_buf = [None]

    def timed_device(d_theta, d_out, n, steps, mdl=None, dd=None):
        """K device-resident passes; returns ms (max over ranks).  Inputs larger than L2: the K passes run back to back between one pair
        of events.  Inputs that would fit in L2 (the strong split at N >= 4): a 256 MB buffer is written between the passes, outside
        the timed regions - every pass is bracketed by its own pair of events on the launching stream and the K durations are summed."""
        mdl, dd = mdl or model, dd or d
        if dd * n * 8 > L2_BYTES:
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                mdl.loglike_ptr(d_theta.data_ptr(), n, n, d_out.data_ptr(), _lib.CC_MEM_DEVICE, stream)
            e1.record()
            barrier()
            return reduce_ranks(e0.elapsed_time(e1))
        if flush_buf[0] is None:
            flush_buf[0] = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        barrier()
        evs = []
        for _ in range(steps):
            flush_buf[0].fill_(1)                      # evicts theta and the work buffers of the previous pass from the 126 MB L2
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            mdl.loglike_ptr(d_theta.data_ptr(), n, n, d_out.data_ptr(), _lib.CC_MEM_DEVICE, stream)
            e1.record()
            evs.append((e0, e1))
        barrier()
        return reduce_ranks(sum(a.elapsed_time(b) for a, b in evs))

    # ---------------------------------------------------------------- strong split of the V vectors (value)
    theta_all = synth.random_theta(spec, V, seed=synth.SEED_THETA, bad_frac=0.01)      # the same V vectors on every rank
    off, Vs = ccdist.shard_vectors(V, world, rank)
    theta = theta_all[off:off + Vs]
    h_theta = torch.from_numpy(np.ascontiguousarray(theta.T)).pin_memory()       # SoA [d][Vs], pinned
    h_out = torch.empty(Vs, dtype=torch.float64).pin_memory()
    d_theta = h_theta.to(dev)
    d_out = torch.empty(Vs, dtype=torch.float64, device=dev)
    for _ in range(args.warmup):
        model.loglike_ptr(d_theta.data_ptr(), Vs, Vs, d_out.data_ptr(), _lib.CC_MEM_DEVICE, stream)
    clocks = ClockSampler(local_rank).start()
    ms_total = timed_device(d_theta, d_out, Vs, args.steps)
    clk = clocks.stop()
    value = V * args.steps / (ms_total * 1e-3)
    # per-launch kernel time (CUDA events on the launching stream, recorded inside cc_loglike_batch)
    kernel_ms = []
    for _ in range(3):
        model.loglike_ptr(d_theta.data_ptr(), Vs, Vs, d_out.data_ptr(), _lib.CC_MEM_DEVICE, stream)
        torch.cuda.synchronize()
        kernel_ms.append(model.last_kernel_ms()[0])
    kern_ms = float(np.mean(kernel_ms))
    launches_per_step = model.last_kernel_ms()[1]
    ll = d_out.cpu().numpy()
    full_frac = float(np.mean(ll != _lib.CC_OVERFLO_LOGLIKE))

    # ---------------------------------------------------------------- weak figure: V vectors on every GPU (N > 1)
    weak = None
    if world > 1:
        thw = synth.random_theta(spec, V, seed=synth.SEED_THETA + rank, bad_frac=0.01) if rank else theta_all
        d_thw = torch.from_numpy(np.ascontiguousarray(thw.T)).to(dev)
        d_outw = torch.empty(V, dtype=torch.float64, device=dev)
        for _ in range(2):
            model.loglike_ptr(d_thw.data_ptr(), V, V, d_outw.data_ptr(), _lib.CC_MEM_DEVICE, stream)
        msw = timed_device(d_thw, d_outw, V, args.steps)
        weak = {"value": V * world * args.steps / (msw * 1e-3), "unit": UNIT, "vectors_per_gpu_per_step": V, "ms_per_step": msw / args.steps}
        del d_thw, d_outw, thw

    # ---------------------------------------------------------------- posterior-like vectors (truth x (1 + 2% noise)), N = 1 only:
    # the regime an MCMC actually visits (3.3 % of the box-uniform draws above have sod < 0.034, where the gamma table falls
    # back to the general incomplete-gamma algorithm, and 1 % are deliberately invalid)
    value_post = None
    if world == 1:
        near = synth.random_theta(spec, V, seed=synth.SEED_THETA + 1000, bad_frac=0.0, near_truth=synth.true_theta(spec, truth), jitter=0.02)
        d_near = torch.from_numpy(np.ascontiguousarray(near.T)).to(dev)
        d_out2 = torch.empty(V, dtype=torch.float64, device=dev)
        for _ in range(2):
            model.loglike_ptr(d_near.data_ptr(), V, V, d_out2.data_ptr(), _lib.CC_MEM_DEVICE, stream)
        value_post = V * args.steps / (timed_device(d_near, d_out2, V, args.steps) * 1e-3)
        del d_near, near, d_out2

    # ---------------------------------------------------------------- the other BASELINE shapes, device-resident (N = 1 only): cfg2
    # (5 x 200, logit-normal serology, seroreversion: the hazard convolution is a second Toeplitz product), cfg4 (17 x 400), and
    # cfg3 / cfg4 with seroreversion ("also report yes")
    other_shapes = None
    if world == 1 and not args.no_other_shapes:
        other_shapes = {}
        for cfg_o, rev_o in (("cfg2", None), ("cfg4", None), ("cfg3", True), ("cfg4", True)):
            spec_o, truth_o = synth.make_spec(cfg_o, serorev=rev_o)
            model_o = _lib.Model(spec_o, device=local_rank)
            Vo = 1 << 18
            th_o = synth.random_theta(spec_o, Vo, seed=synth.SEED_THETA + 7, bad_frac=0.01)
            d_tho = torch.from_numpy(np.ascontiguousarray(th_o.T)).to(dev)
            d_outo = torch.empty(Vo, dtype=torch.float64, device=dev)
            for _ in range(2):
                model_o.loglike_ptr(d_tho.data_ptr(), Vo, Vo, d_outo.data_ptr(), _lib.CC_MEM_DEVICE, stream)
            rate_o = Vo * args.steps / (timed_device(d_tho, d_outo, Vo, args.steps, model_o, spec_o.d) * 1e-3)
            full_o = float(np.mean(d_outo.cpu().numpy() != _lib.CC_OVERFLO_LOGLIKE))
            other_shapes[cfg_o + ("_serorev" if rev_o else "")] = {"value": rate_o, "unit": UNIT, "vectors_per_step": Vo, "d": spec_o.d, "flops_per_eval": f_alg(spec_o),
                                   "full_eval_fraction": full_o, "_falg_rate": rate_o * f_alg(spec_o) * full_o / 1e12}
            del d_tho, d_outo, th_o
            model_o.close()

    # ---------------------------------------------------------------- end-to-end leg (host buffers through the C ABI)
    def timed_host(theta_ptr, out_ptr, steps):
        barrier()
        wall0 = time.perf_counter()
        for _ in range(steps):
            model.loglike_ptr(theta_ptr, Vs, Vs, out_ptr, _lib.CC_MEM_HOST, 0)
        torch.cuda.synchronize()
        dt = time.perf_counter() - wall0
        barrier()
        return reduce_ranks(dt)

    e2e_steps = max(2, args.steps // 2)
    timed_host(h_theta.data_ptr(), h_out.data_ptr(), max(1, args.warmup // 2))
    e2e_value = V * e2e_steps / timed_host(h_theta.data_ptr(), h_out.data_ptr(), e2e_steps)
    assert np.array_equal(h_out.numpy(), ll), "host path and device path must agree bit for bit"
    # the same call from PAGEABLE memory (plain numpy arrays, as an R caller's REAL(theta) would be): the library stages it
    # through its own pinned ring
    p_theta = np.ascontiguousarray(theta.T).copy()
    p_out = np.empty(Vs)
    timed_host(p_theta.ctypes.data, p_out.ctypes.data, 1)
    e2e_pageable = V * e2e_steps / timed_host(p_theta.ctypes.data, p_out.ctypes.data, e2e_steps)
    assert np.array_equal(p_out, ll), "pageable host path and device path must agree bit for bit"
    del p_theta
    # the PCIe ceiling of that leg: the same pinned buffer copied alone (reported, not part of any timed region above)
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d_theta.copy_(h_theta, non_blocking=True)
    c0.record()
    for _ in range(3):
        d_theta.copy_(h_theta, non_blocking=True)
    c1.record()
    torch.cuda.synchronize()
    h2d_gbs = 3 * d * Vs * 8 / (c0.elapsed_time(c1) * 1e-3) / 1e9

    # ---------------------------------------------------------------- MCMC legs
    def oracle_check(spec_, out, n_max):
        """final cold-rung states re-evaluated by the CPU oracle (rank 0): max relative error of the recorded log-likelihoods"""
        from oracle.oracle import Oracle
        th = out["theta"][:n_max, 0, -1, :]
        chk = Oracle(spec_).loglike(th, nthreads=min(host_threads(), 16))
        got = out["loglike"][:n_max, 0, -1]
        bad = (chk == _lib.CC_OVERFLO_LOGLIKE) | (got == _lib.CC_OVERFLO_LOGLIKE)
        rel = np.abs(got[~bad] - chk[~bad]) / np.abs(chk[~bad])
        return {"states": int(len(th)), "max_rel_err_vs_oracle": float(rel.max()) if rel.size else 0.0,
                "failure_value_mismatches": int(np.sum(bad & (chk != got)))}

    mcmc = mcmc2 = mcmc3 = collective = None
    if not args.no_mcmc:
        # ---- cfg3: 64 chains x 50 rungs on ONE GPU (BASELINE configs[2]); every rank runs its own 64 chains when N > 1
        chains3, rungs3 = truth["chains"], truth["rungs"]
        init3 = np.tile(synth.true_theta(spec, truth), (chains3, 1))
        settle3, probe3 = 100, 50
        cap3 = 6000
        mc = _lib.Mcmc(model, burnin=settle3 + probe3 + cap3 + 2, samples=0, rungs=rungs3, chains=chains3, chain_offset=rank * chains3,
                       seed=synth.SEED_MCMC, record_rungs=False, init=init3, thin=50)
        mc.advance(settle3)               # untimed: lets the Robbins-Monro bandwidths settle (bw starts at 1)
        mc.advance(probe3)
        it3 = int(min(cap3, max(100, args.cfg3_seconds * 1e3 / (mc.last_advance_ms / probe3))))
        clocks = ClockSampler(local_rank).start()
        barrier()
        mc.advance(it3)
        ms3 = reduce_ranks(mc.last_advance_ms)
        clk3 = clocks.stop()
        barrier()
        P3 = chains3 * rungs3 * world
        mcmc3 = {"iter_per_sec_all_chains_x_rungs": P3 * it3 / (ms3 * 1e-3), "chains_per_gpu": chains3, "rungs": rungs3,
                 "iterations_timed": it3, "device_seconds": ms3 * 1e-3, "loglike_evals_per_sec": P3 * it3 * d / (ms3 * 1e-3),
                 "ms_per_iteration": ms3 / it3, "settle_iterations": settle3 + probe3, "sm_mhz": clk3.get("sm_mhz"),
                 "sm_min_mhz_over_ranks": reduce_ranks(clk3.get("sm_min_mhz") or 0.0, "min"), "reasons": clk3.get("reasons"),
                 "timing": "CUDA events on the library's stream around the %d iterations (2 launches each), max over ranks" % it3,
                 "workload": "cfg3: 64 chains x 50 rungs per GPU, d=34 (one iteration = d single-site updates + coupling pass)"}
        if rank == 0:
            mcmc3["check"] = oracle_check(spec, mc.fetch(), 64)
        mc.close()

        # ---- cfg2: 4 chains x 20 rungs (BASELINE configs[1]; 80 particles: a tenth of one wave of warps - reported for completeness)
        mcmc2 = None
        if world == 1:
            spec2, truth2 = synth.make_spec("cfg2")
            model2 = _lib.Model(spec2, device=local_rank)
            init2 = np.tile(synth.true_theta(spec2, truth2), (truth2["chains"], 1))
            it2 = 1500
            mc2 = _lib.Mcmc(model2, burnin=100 + it2 + 2, samples=0, rungs=truth2["rungs"], chains=truth2["chains"], seed=synth.SEED_MCMC,
                            record_rungs=False, init=init2, thin=50)
            mc2.advance(100)
            mc2.advance(it2)
            P2 = truth2["chains"] * truth2["rungs"]
            mcmc2 = {"iter_per_sec_all_chains_x_rungs": P2 * it2 / (mc2.last_advance_ms * 1e-3), "chains": truth2["chains"], "rungs": truth2["rungs"],
                     "iterations_timed": it2, "ms_per_iteration": mc2.last_advance_ms / it2,
                     "loglike_evals_per_sec": P2 * it2 * spec2.d / (mc2.last_advance_ms * 1e-3),
                     "workload": "cfg2: 5 age bands x 200 days, logit-normal serology, seroreversion, d=%d, 4 chains x 20 rungs" % spec2.d,
                     "check": oracle_check(spec2, mc2.fetch(), 4)}
            mc2.close()
            model2.close()

        # ---- cfg4: 4096 chains x 64 rungs sharded over the ranks (BASELINE configs[3])
        spec4, truth4 = synth.make_spec("cfg4")
        model4 = _lib.Model(spec4, device=local_rank)
        chains4 = args.mcmc_chains or truth4["chains"]
        rungs4 = truth4["rungs"]
        c_off, c_loc = ccdist.shard_chains(chains4, world, rank)
        it4 = args.mcmc_iters if args.mcmc_iters > 0 else int(min(200, max(12, 25 * world)))
        settle4 = args.mcmc_settle
        init4 = np.tile(synth.true_theta(spec4, truth4), (c_loc, 1))
        # the timed iterations are the sampling phase (their cold-rung draws are what the collective gathers)
        mc4 = _lib.Mcmc(model4, burnin=settle4 + 1, samples=it4, rungs=rungs4, chains=c_loc, chain_offset=c_off, seed=synth.SEED_MCMC,
                        record_rungs=False, init=init4, GTI_pow=3.0)
        mc4.advance(settle4)
        clocks = ClockSampler(local_rank).start()
        barrier()
        mc4.advance(it4)
        ms4 = reduce_ranks(mc4.last_advance_ms)
        clk4 = clocks.stop()
        barrier()
        P4 = chains4 * rungs4
        mcmc = {"iter_per_sec_all_chains_x_rungs": P4 * it4 / (ms4 * 1e-3), "chains": chains4, "rungs": rungs4, "chains_per_gpu": c_loc,
                "iterations_timed": it4, "device_seconds": ms4 * 1e-3, "loglike_evals_per_sec": P4 * it4 * spec4.d / (ms4 * 1e-3),
                "ms_per_iteration": ms4 / it4, "settle_iterations": settle4, "scaling": "strong (chains sharded, no per-iteration communication)",
                "sm_mhz": clk4.get("sm_mhz"), "sm_min_mhz_over_ranks": reduce_ranks(clk4.get("sm_min_mhz") or 0.0, "min"),
                "reasons": clk4.get("reasons"), "power_w_max": clk4.get("power_w_max"),
                "workload": "cfg4: 17 age bands x 400 days, d=48, %d chains x %d rungs sharded over %d GPU(s)" % (chains4, rungs4, world)}
        out4 = mc4.fetch() if rank == 0 else None
        if rank == 0:
            mcmc["check"] = oracle_check(spec4, out4, 32)

        # ---- the run's only collective: all-gather of the cold-rung SAMPLING draws (device tensors, NCCL) + R-hat
        rows = mc4.rows
        draws = ccdist.draws_tensor(mc4, dev)[:, 0, rows - it4:rows, :].contiguous()      # [chains_local, it4, d] in HBM
        torch.cuda.synchronize()
        gathered = ccdist.allgather_draws(draws, chains4)                                     # warm-up (NCCL connection set-up)
        rh_red = ccdist.rhat_allreduce(draws, chains4)
        reps, ag_ms, ar_ms = 5, [], []
        for _ in range(reps):
            barrier()
            g0, g1, g2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            g0.record()
            gathered = ccdist.allgather_draws(draws, chains4)
            g1.record()
            rh_red = ccdist.rhat_allreduce(draws, chains4)
            g2.record()
            torch.cuda.synchronize()
            ag_ms.append(reduce_ranks(g0.elapsed_time(g1)))
            ar_ms.append(reduce_ranks(g1.elapsed_time(g2)))
        rh_gat = ccdist.rhat_torch(gathered)
        total_bytes = gathered.numel() * 8
        ag = float(np.median(ag_ms))
        collective = {
            "backend": "nccl" if world > 1 else "none (single process: the gather is the identity)", "world": world,
            "allgather": {"what": "cold-rung sampling draws [chains, iterations, d] f64, device tensors (zero-copy view of the run state)",
                          "shape": list(gathered.shape), "bytes_total": total_bytes, "bytes_per_rank": draws.numel() * 8, "ms": ag,
                          "algbw_GBps": total_bytes / (ag * 1e-3) / 1e9, "busbw_GBps": total_bytes * (world - 1) / world / (ag * 1e-3) / 1e9},
            "rhat_allreduce": {"what": "3 x d per-chain moment sums, one all-reduce (includes the local reductions over the draws)",
                               "bytes": 3 * spec4.d * 8, "ms": float(np.median(ar_ms))},
            "rhat_max_abs_diff_allreduce_vs_gathered": float((rh_red - rh_gat).abs().max().item()),
            "rhat_median": float(rh_gat.median().item()), "timing": "CUDA events, median of %d, max over ranks" % reps}
        # ---- posterior summaries ON THE DEVICE next to the gathered draws (csrc/cc_summ.cu; R/summarize_plot.R:46-153):
        # (a) coda::gelman.diag + drjacoby R-hat without moving a draw: per-chain moments by the library's kernel, an NCCL
        #     all-gather of 2 x chains x d numbers, the finishing kernel; (b) the credible-interval table pooled over all 4096
        #     chains from the gathered tensor (radix-select quantiles, AR spectrum -> ESS / Geweke); (c) the same per chain,
        #     every rank summarising its own chains, only the rows travelling
        from covidcurve_b200 import summaries as ccsum
        def timed_ms(fn, reps=3):
            ts = []
            for _ in range(reps):
                barrier()
                t0 = time.perf_counter()
                r = fn()
                torch.cuda.synchronize()
                ts.append(reduce_ranks((time.perf_counter() - t0) * 1e3))
            return r, float(np.median(ts))
        for _ in range(2):                                                                    # warm-up (imports, NCCL set-up)
            ccdist.gelman_allgather(draws, chains4)
        (psrf, psrf_up, rh_k), ms_gel = timed_ms(lambda: ccdist.gelman_allgather(draws, chains4))
        pooled, ms_pool = timed_ms(lambda: _lib.summarize_draws(gathered, by_chain=False))
        bychain, ms_chain = timed_ms(lambda: ccdist.cred_intervals_allgather(draws, chains4, spec4.names))
        collective["device_summaries"] = {
            "gelman_allgather": {"what": "cc_chain_moments on local draws -> all-gather of [chains, 2, d] moments -> cc_gelman_from_moments",
                                 "bytes_gathered": 2 * chains4 * spec4.d * 8, "ms": ms_gel, "psrf_median": float(np.median(psrf)),
                                 "psrf_upper_median": float(np.median(psrf_up))},
            "pooled_table": {"what": "cc_summarize_draws(by_chain=0) on the gathered tensor: %d series of %d draws" % (spec4.d, chains4 * it4),
                             "ms": ms_pool},
            "by_chain_table": {"what": "cc_summarize_draws(by_chain=1) per rank on its own chains + all-gather of the rows",
                               "rows": int(len(bychain)), "ms": ms_chain},
            "timing": "host wall clock around the call incl. synchronisation, max over ranks, median of 3"}
        if rank == 0:
            from covidcurve_b200.summaries import rhat_per_parameter
            gat = gathered.cpu().numpy()
            half = gat[:, it4 // 2:, :]
            e_host, u_host = ccsum.gelman_diag(gat)
            q_host = np.quantile(gat.reshape(-1, spec4.d), [0.025, 0.5, 0.975], axis=0).T
            ds = collective["device_summaries"]
            ds["gelman_allgather"]["max_rel_diff_vs_host_numpy"] = float(np.nanmax(np.abs(psrf - e_host) / e_host))
            ds["gelman_allgather"]["rhat_max_abs_diff_vs_host_numpy"] = float(np.nanmax(np.abs(rh_k - rhat_per_parameter(half))))
            ds["pooled_table"]["quantiles_equal_numpy"] = bool(np.array_equal(pooled[0][:, [1, 2, 4]], q_host))
            ds["by_chain_table"]["medians_equal_numpy"] = bool(np.array_equal(
                bychain["median"].to_numpy().reshape(chains4, spec4.d), np.quantile(gat, 0.5, axis=1)))
            host = rhat_per_parameter(gat)
            collective["rhat_max_abs_diff_device_vs_host_numpy"] = float(np.max(np.abs(host - rh_gat.cpu().numpy())))
            # the gathered block of rank 0's own chains is bit for bit what rank 0 recorded
            collective["gathered_equals_local"] = bool(np.array_equal(gathered[:c_loc].cpu().numpy(), out4["theta"][:, 0, rows - it4:rows, :]))
        del gathered, draws
        mc4.close()
        model4.close()

    # ---------------------------------------------------------------- roofline + CPU baseline (rank 0)
    if rank == 0:
        dfma_tf, _ = _lib.fp64_peak_probe(local_rank, 8192)
        dmma_tf, _, _ = _lib.dmma_probe(local_rank, 8192, 0)
        peak_tf = max(dfma_tf, dmma_tf)
        falg = f_alg(spec)
        alg_bytes = (8 * d + 8) * Vs
        achieved = falg * Vs * full_frac / (kern_ms * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        traffic, traffic_src = measured_traffic()
        roofline = {
            "bound": "fp64", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
            "traffic": traffic["bytes_per_eval"] * Vs if traffic else None, "traffic_source": traffic_src,
            # hardware view of the same kernel from the same capture: the FP64 vector and tensor sub-pipes share one pipe
            "ncu": (traffic or {}).get("ncu"),
            "algorithmic_bytes": float(alg_bytes),
            "kernel": "loglike_phase2_kernel (the per-day pass; kernel_ms covers the whole pipeline of one cc_loglike_batch call)",
            "kernel_ms": kern_ms, "vectors_per_launch": Vs, "flops_per_eval": falg, "full_eval_fraction": full_frac,
            "peak_source": "max of cc_fp64_peak_probe (register-resident DFMA chains: %.2f) and cc_dmma_probe (mma.sync m8n8k4 f64: %.2f), measured "
                           "live in this run; MEASURED_PEAKS.json has no FP64 entry" % (dfma_tf, dmma_tf),
            "hbm": {"achieved": alg_bytes / (kern_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": alg_bytes / (kern_ms * 1e-3) / 1e9 / hbm_peak,
                    "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"},
        }
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle.oracle import Oracle
            o = Oracle(spec)
            cores = host_threads()
            rate, n, dt, chk = cpu_loglike_rate(o, theta, args.cpu_seconds, cores)
            rate1, n1, dt1, _ = cpu_loglike_rate(o, theta, 4.0, 1)
            bad = (chk == _lib.CC_OVERFLO_LOGLIKE) | (ll[:n] == _lib.CC_OVERFLO_LOGLIKE)
            relv = np.zeros(n)
            relv[~bad] = np.abs(chk[~bad] - ll[:n][~bad]) / np.abs(chk[~bad])
            mism = bad & (chk != ll[:n])
            worst = np.argsort(relv)[-3:][::-1]
            cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "first %d of the step's vectors, %.1f s, %d threads (g++ -O2 restatement of the reference; the reference "
                             "itself needs R/Rcpp/libRmath, absent here)" % (n, dt, cores),
                   "value_1_core": rate1, "sample_1_core": "first %d vectors, %.1f s, 1 thread" % (n1, dt1),
                   "max_rel_err_gpu_vs_cpu": float(relv.max()), "failure_value_mismatches": int(mism.sum()),
                   "p999_rel_err_gpu_vs_cpu": float(np.quantile(relv, 0.999)), "vectors_above_1e-10": int(np.sum(relv > 1e-10)),
                   "worst_vectors": [{"index": int(i), "rel_err": float(relv[i]), "sod": float(theta[i, spec.index("sod")]),
                                      "mod": float(theta[i, spec.index("mod")])} for i in worst]}
            if not args.no_mcmc:
                cpu["mcmc"] = {"cfg3_all_cores": cpu_mcmc_rate("cfg3", cores, 6.0), "cfg3_1_core": cpu_mcmc_rate("cfg3", 1, 5.0),
                               "cfg4_all_cores": cpu_mcmc_rate("cfg4", cores, 8.0), "cfg2_all_cores": cpu_mcmc_rate("cfg2", cores, 4.0),
                               "note": "oracle port of drjacoby's loop (one chain per thread, 4 rungs): chain-rung iterations/s, the unit of `mcmc`"}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "vectors_per_step": V, "vectors_per_gpu_per_step": Vs, "layout": "SoA theta[d][V]",
                       "l2": "inputs larger than L2 (%.0f MB per GPU per step), passes back to back" % (d * Vs * 8 / 1e6) if d * Vs * 8 > 126e6 else
                             "inputs of %.0f MB per GPU per step would fit in L2: L2 flushed between the timed passes (256 MB written, outside "
                             "the per-pass event pairs whose durations are summed)" % (d * Vs * 8 / 1e6),
                       "full_eval_fraction": full_frac, "split": "V vectors split evenly over the ranks (SURVEY.md 8e), no collective"},
            "weak": weak,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": d * V * 8, "d2h_bytes_per_step": V * 8, "steps": e2e_steps,
                    "h2d_gbs_measured_per_rank": h2d_gbs, "h2d_bound": h2d_gbs * 1e9 / (d * 8) * world, "cpu_affinity": numa,
                    "pageable": {"value": e2e_pageable, "unit": UNIT,
                                 "note": "caller buffers in pageable memory, staged by the library through its pinned ring"}},
            "gpu_launches": args.steps * launches_per_step,
            "posterior_like": None if value_post is None else {"value": value_post, "unit": UNIT, "frac_of_fp64_peak": value_post * falg / 1e12 / peak_tf,
                                                                "vectors": "truth x (1 + 0.02 N(0,1)), clipped to the box"},
            "other_shapes": None if other_shapes is None else {k: dict({kk: vv for kk, vv in v.items() if kk != "_falg_rate"},
                                                                       frac_of_fp64_peak=v["_falg_rate"] / peak_tf) for k, v in other_shapes.items()},
            "clocks": clk, "roofline": roofline, "cpu_baseline": cpu, "mcmc": mcmc, "mcmc_cfg3": mcmc3, "mcmc_cfg2": mcmc2, "collective": collective,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
