#!/usr/bin/env python
"""bench.py -- throughput of the LongTR read x haplotype hot path on B200.

One "step" = one pass of the hot path (Viterbi log-likelihoods of every pooled read against
every candidate haplotype + genotype posteriors) over one batch of synthetic loci
(BASELINE.json config 3: HiFi STRs, generator of SURVEY.md section 8d).  Loci are sharded over
the ranks with no data-path collective (weak scaling: every GPU gets --loci loci).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--loci L] [--config 3|4]
  python bench.py --impl reference ...     # the reference's own CPU code on the host cores

Prints ONE JSON line (rank 0).  value = loci/s with inputs resident in HBM (plan kernels included every step); e2e = the
same through ltr_job_submit / ltr_job_wait with pinned HOST buffers on ONE host thread (H2D + plan + kernels + D2H inside
the timed region, up to three jobs in flight).  The default run appends extra.c4 / extra.c5: BASELINE.json configs[3] and
configs[4] timed the same way.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

FP64_OPS_PER_CELL = 17.0  # 10 DADD + 7 compare/select per DP cell


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=[3, 4, 5])
    ap.add_argument("--n2", action="store_true", help="only the extra.n2 line (clustering kernels)")
    ap.add_argument("--n3", action="store_true", help="only the extra.n3 line (BAM files -> calls)")
    ap.add_argument("--em", action="store_true", help="only the extra.em line (stutter-model EM kernel)")
    ap.add_argument("--one-process-devices", type=int, default=0,
                    help="only the raw-loci arm (ltr_genotyper_run) with ONE process driving this many devices")
    ap.add_argument("--loci", type=int, default=0, help="loci per GPU per step (default: config size)")
    ap.add_argument("--cpu-sample-loci", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra.c4 / extra.c5 sub-lines of the default run")
    ap.add_argument("--no-raw", action="store_true", help="skip the e2e_from_flat_loci arm (raw loci -> calls)")
    return ap.parse_args()


CONFIG_LOCI = {3: 100000, 4: 10000, 5: 50000}
CONFIG_NAME = {3: "synthetic HiFi STRs: 1-6 bp motifs, 50-300 bp repeats, 30 reads/locus (BASELINE.json configs[2])",
               4: "synthetic VNTRs: 500-1000 bp repeats, 2-12 haplotypes, ONT-like params (BASELINE.json configs[3])",
               5: "homopolymer stutter path: --stutter-align-len 20, synthetic homopolymer loci (BASELINE.json configs[4])"}
FP64_OPS_PER_CELL_SHORT = 13.0  # flank-row cell of align_seq_to_hap_short: 9 add + 4 max (HapAligner.cpp:141-156)


class ClockSampler:
    """Samples SM clocks / throttle reasons during the timed region (B200_PROFILING.md): NVML every 5 ms when the
    binding is available (the timed region of config 3 is only ~0.2 s long), else `nvidia-smi -lms 100`."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    # nvmlClocksThrottleReason* / nvmlClocksEventReason* bits
    REASON_BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None
        self.nvml = None
        self.samples = []  # (sm_mhz, reasons bitmask) from NVML
        self.max_mhz = None
        self.stop_flag = False
        self.thread = None

    def _nvml_open(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            self.nvml = (pynvml, h)
            return True
        except Exception:
            self.nvml = None
            return False

    def _nvml_loop(self):
        pynvml, h = self.nvml
        reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(pynvml, "nvmlDeviceGetCurrentClocksThrottleReasons", None)
        while not self.stop_flag:
            try:
                mhz = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                bits = int(reasons_fn(h)) if reasons_fn else 0
                self.samples.append((mhz, bits))
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self._nvml_open():
            self.thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            sm = [s[0] for s in self.samples]
            bits = 0
            for s in self.samples:
                bits |= s[1]
            reasons = sorted(nm for nm, b in self.REASON_BITS.items() if bits & b)
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "samples": len(sm),
                    "reasons": reasons, "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi"}


NUMA_CPUS = None  # CPUs this rank was bound to (None: not bound)


def bind_to_gpu_numa_node(local):
    """With several ranks on one box every rank pins ~0.75 GB of host buffers per step for its GPU; bound to the CPUs next to
    that GPU (NVML's affinity mask) the buffers are first touched -- and therefore placed -- on the GPU's own NUMA node, and
    the copies of eight ranks do not cross the socket interconnect.  What `numactl --cpunodebind` would do for a production
    launch; LTR_BENCH_NO_BIND=1 switches it off."""
    global NUMA_CPUS
    if os.environ.get("LTR_BENCH_NO_BIND"):
        return
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus and len(cpus) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, cpus)
            NUMA_CPUS = len(cpus)
    except Exception:
        pass


def dist_setup(n_gpus):
    """One process per GPU; NCCL only for the barrier and the max-over-ranks of the timings."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        bind_to_gpu_numa_node(local)
    import torch
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    return torch, rank, world, local


def barrier(torch, world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def ncu_traffic(key):
    """DRAM bytes of the dominant kernel from the committed `ncu --set full` capture (profiles/traffic.json)."""
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "traffic.json")) as f:
            return json.load(f).get(key)
    except OSError:
        return None


def max_over_ranks(torch, world, x):
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(torch, world, x):
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def pinned_copy(torch, d):
    """Copy a dict of numpy arrays into pinned host memory (e2e inputs come from pinned buffers)."""
    out, keep = {}, []
    for k, v in d.items():
        if v is None:
            out[k] = None
            continue
        t = torch.empty(max(1, v.nbytes), dtype=torch.uint8, pin_memory=True)
        a = t.numpy()[:v.nbytes].view(v.dtype)
        a[...] = v
        out[k] = a
        keep.append(t)
    return out, keep


def cpu_baseline(work, n_sample, threads):
    """The reference's CPU path (process_reads + posteriors) on a bounded sample of the same workload.
    oracle/ is used here only as the measured CPU baseline, never on the product path.
    Returns (kind, seconds inside the hot path [max over threads], wall seconds)."""
    from oracle import pyoracle as po
    sb, sp = work.subset(n_sample)
    t0 = time.perf_counter()
    if po.ref_available():
        kind = "reference"
        _ll, sec, _post = po.ref_viterbi_batch(sb, work.aln_params, n_threads=threads, post=sp)
    else:
        kind = "port"
        ll, _cells = po.viterbi_batch(sb, aln_params=work.aln_params, n_threads=threads)
        H = np.diff(sb["locus_hap_begin"]).astype(np.int64)
        P = np.diff(sb["locus_read_begin"]).astype(np.int64)
        off, lsb = 0, sp["locus_sread_begin"]
        for l in range(n_sample):
            h, p = int(H[l]), int(P[l])
            mat = ll[off:off + h * p].reshape(p, h)
            off += h * p
            r0, r1 = int(lsb[l]), int(lsb[l + 1])
            po.log_sample_posteriors(mat[sp["pool_index"][r0:r1]], sp["log_p1"][r0:r1], sp["log_p2"][r0:r1],
                                     sp["sample_label"][r0:r1], 1)
        sec = time.perf_counter() - t0
    wall = time.perf_counter() - t0
    return kind, sec, wall


def sample_cells(work, n_sample):
    b = work.batch
    hoff, roff = b["hap_off"].astype(np.int64), b["read_off"].astype(np.int64)
    lhb, lrb = b["locus_hap_begin"], b["locus_read_begin"]
    cells = 0
    for l in range(n_sample):
        n = np.diff(hoff[lhb[l]:lhb[l + 1] + 1]) - 60
        m = np.diff(roff[lrb[l]:lrb[l + 1] + 1])
        nm = n[:, None] * m[None, :]
        nm[np.abs(n[:, None] - m[None, :]) > 600] = 0
        cells += int(nm[n > 0].sum())
    return cells


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from longtr_b200 import workloads
    threads = os.cpu_count() or 1
    n_sample = args.cpu_sample_loci or (100 * threads if args.config == 3 else max(2, threads))
    work = workloads.generate(args.config, n_sample)
    cells = sample_cells(work, n_sample)
    times = []
    kind = "port"
    for it in range(args.warmup + args.steps):
        kind, sec, wall = cpu_baseline(work, n_sample, threads)
        if it >= args.warmup:
            times.append(wall)
    tot = sum(times)
    val = n_sample * len(times) / tot
    line = {
        "impl": "reference", "metric": "loci_per_sec", "value": val, "unit": "loci/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "gcups": cells * len(times) / tot / 1e9,
        "config": {"workload": CONFIG_NAME[args.config], "loci_per_step": n_sample,
                   "note": "bounded sample of the b200 arm's workload (same generator and seeds)"},
        "cpu_baseline": {"value": val, "unit": "loci/s", "cores": threads, "kind": kind,
                         "sample": "%d loci per step, wall clock incl. Haplotype construction" % n_sample},
        "e2e": {"value": val, "unit": "loci/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_stutter(work, n_sample, threads):
    """Reference HapAligner::process_reads (homopolymer path) over the first n_sample loci, one locus per task on a
    thread pool (ctypes releases the GIL; the README's 'split the BED' parallelisation).  Returns wall seconds."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import pyoracle as po
    which = "ref" if po.ref_available() else "oracle"
    loci = [work.flat_locus(l) for l in range(n_sample)]
    po.process_reads(loci[0][0][0], loci[0][1][0], loci[0][1][1], which=which)  # one-time table init outside the clock
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(lambda x: po.process_reads(x[0][0], x[1][0], x[1][1], which=which), loci))
    return ("reference" if which == "ref" else "port"), time.perf_counter() - t0


def run_reference_stutter(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from longtr_b200 import workloads
    threads = os.cpu_count() or 1
    n_sample = args.cpu_sample_loci or 8 * threads
    work = workloads.generate_stutter(n_sample)
    cells = work.cells(n_sample)
    times, kind = [], "port"
    for it in range(args.warmup + args.steps):
        kind, wall = cpu_baseline_stutter(work, n_sample, threads)
        if it >= args.warmup:
            times.append(wall)
    tot = sum(times)
    val = n_sample * len(times) / tot
    print(json.dumps({
        "impl": "reference", "metric": "loci_per_sec", "value": val, "unit": "loci/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "gcups": cells * len(times) / tot / 1e9,
        "config": {"workload": CONFIG_NAME[5], "loci_per_step": n_sample,
                   "note": "bounded sample of the b200 arm's workload (same generator and seeds)"},
        "cpu_baseline": {"value": val, "unit": "loci/s", "cores": threads, "kind": kind,
                         "sample": "%d loci per step, wall clock incl. Haplotype construction" % n_sample},
        "e2e": {"value": val, "unit": "loci/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}), flush=True)


def run_stutter(args, torch, rank, world, local, eng, n_loci, steps, warmup, with_cpu_baseline):
    """--config 5: the homopolymer / --stutter-align-len path (kernel 2) at the process_reads boundary."""
    from longtr_b200 import workloads
    fp64_rate = max(eng.fp64_issue_rate(0)[0] for _ in range(2))
    work = workloads.generate_stutter(n_loci, first_locus=rank * n_loci)
    pinned_b, keep_b = pinned_copy(torch, work.batch)
    out = np.zeros(abi_ll_size(work.batch), dtype=np.float64)
    for _ in range(warmup):
        eng.stutter_ll(pinned_b, out=out)
    sampler = ClockSampler(local)
    barrier(torch, world)
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    dev_ms, launches = 0.0, 0
    for _ in range(steps):
        out, st = eng.stutter_ll(pinned_b, out=out)
        dev_ms += st.kernel_ms
        launches += st.n_launches
    barrier(torch, world)
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if rank == 0 else None
    step_ms = max_over_ranks(torch, world, dev_ms / steps)
    e2e_ms = max_over_ranks(torch, world, wall_ms / steps)
    total_loci = sum_over_ranks(torch, world, float(n_loci))
    total_cells = sum_over_ranks(torch, world, float(st.n_cells))
    if rank != 0:
        return None
    achieved = st.n_cells / (dev_ms / steps) / 1e6
    peak = fp64_rate / FP64_OPS_PER_CELL_SHORT / 1e9
    line = {
        "metric": "loci_per_sec", "value": total_loci / (step_ms * 1e-3), "unit": "loci/s", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": step_ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "gcups": total_cells / (step_ms * 1e-3) / 1e9,
        "config": {"workload": CONFIG_NAME[5], "loci_per_gpu": n_loci, "pairs_per_gpu": int(st.n_pairs),
                   "cell_equivalents_per_gpu": int(st.n_cells),
                   "l2": "inputs (%.0f MB) streamed once per step; per-warp working set lives in shared memory" %
                         (work.input_bytes / 1e6),
                   "parallelism": "locus-sharded, no collective",
                   "rank_bound_to_cpus_of_its_gpu": NUMA_CPUS, "ll_checksum": float(np.sum(out))},
        "e2e": {"value": total_loci / (e2e_ms * 1e-3), "unit": "loci/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(st.h2d_bytes), "d2h_bytes_per_step": int(st.d2h_bytes)},
        "gpu_launches": launches, "clocks": clocks,
        "roofline": {"bound": "fp64_issue", "achieved": achieved, "peak": peak, "unit": "GCUPS", "frac": achieved / peak,
                     "traffic": None,
                     "peak_source": "ltr_fp64_issue_rate (DADD lane-ops/s measured in this run) / 13 FP64 ops per "
                                    "flank cell; cell-equivalents as defined in SURVEY.md 8d",
                     "fp64_lane_ops_per_s": fp64_rate}}
    if with_cpu_baseline:
        threads = os.cpu_count() or 1
        n_sample = args.cpu_sample_loci or min(n_loci, 8 * threads)
        kind, wall = cpu_baseline_stutter(work, n_sample, threads)
        line["cpu_baseline"] = {"value": n_sample / wall, "unit": "loci/s", "cores": threads, "kind": kind,
                                "gcups": work.cells(n_sample) / wall / 1e9,
                                "sample": "first %d loci of the same workload, all host threads" % n_sample}
    work.close()
    return line


def abi_ll_size(batch):
    from longtr_b200 import abi
    return abi.stutter_ll_size(batch)


def bench_sample_parity(work, n_sample, ref_ll, ref_post, ll, post):
    """The reference's results on the CPU-baseline sample (the first n_sample loci of the timed workload) against what the
    GPU job produced for the same loci: log-likelihoods bit for bit, posteriors to 1e-12 relative (CUDA exp/log vs libm)."""
    b, p = work.batch, work.post
    H = np.diff(b["locus_hap_begin"][:n_sample + 1]).astype(np.int64)
    n_post = int(np.sum(p["locus_n_samples"][:n_sample].astype(np.int64) * H * H))
    out = {"loci": int(n_sample), "pairs": int(len(ref_ll)),
           "ll_bit_exact": bool(np.array_equal(ref_ll, ll[:len(ref_ll)]))}
    if ref_post is not None and post is not None and n_post == len(ref_post):
        g = post[:n_post]
        ok = np.isfinite(ref_post) & (np.abs(ref_post) > 1e-300)
        rel = float(np.max(np.abs(g[ok] - ref_post[ok]) / np.abs(ref_post[ok]))) if ok.any() else 0.0
        out["post_max_rel_err"] = rel
        out["post_within_1e-12"] = bool(rel <= 1e-12)
    return out


def measure_long(args, config, n_loci, steps, warmup, torch, rank, world, local, eng, peak_gcups, fp64_rate,
                 with_cpu_baseline):
    """Configs 3 / 4 (long path): resident arm (value), asynchronous end-to-end arm (e2e), roofline, CPU baseline."""
    import collections
    from longtr_b200 import abi, workloads
    work = workloads.generate(config, n_loci, first_locus=rank * n_loci)
    pinned_b, keep_b = pinned_copy(torch, work.batch)
    pinned_p, keep_p = pinned_copy(torch, work.post)
    job = eng.create_job(pinned_b, pinned_p, aln_params=work.aln_params)

    # ---- resident arm: inputs in HBM; every step = plan kernels + Viterbi kernels + fan-out + posteriors -------------
    for _ in range(warmup):
        job.run()
    sampler = ClockSampler(local)
    barrier(torch, world)
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    dev_ms = vit_ms = plan_ms = 0.0
    launches = 0
    for _ in range(steps):
        st = job.run()
        dev_ms += st.kernel_ms
        vit_ms += st.viterbi_ms
        plan_ms += st.plan_ms
        launches += st.n_launches
    barrier(torch, world)
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if rank == 0 else None
    st = job.stats()
    step_ms = max_over_ranks(torch, world, max(dev_ms, 0.0) / steps)
    wall_step_ms = max_over_ranks(torch, world, wall_ms / steps)
    total_loci = sum_over_ranks(torch, world, float(n_loci))
    total_cells = sum_over_ranks(torch, world, float(st.n_cells))
    vit_step_ms = max_over_ranks(torch, world, vit_ms / steps)
    ll, post, tot = job.download()
    checksum = float(np.sum(ll[ll > -600.0]))
    n_ll, n_post, n_tot = job.n_ll, job.n_post, job.n_totals
    job.close()

    # ---- end-to-end arm: ONE host thread, host buffers through ltr_job_submit / ltr_job_wait ---------------------------
    # Every step moves its inputs from pinned host memory to the device and its results back inside the timed region; with
    # `depth` jobs in flight the upload of step k+1 and the download of step k-1 overlap the kernels of step k.
    vb, keep_vb = abi.make_viterbi_batch(pinned_b)
    pb, keep_pb = abi.make_posterior_batch(pinned_p)
    prepared = (vb, pb, (keep_vb, keep_pb))
    # the same batch with its reads as ONE 4-bit stream (BAM's own sequence encoding, ltr_ctx_set_read_encoding): half the
    # bytes of the largest copy.  Packed once, outside the clock: this is the form a caller that reads BAM records holds.
    n_bases = int(np.asarray(work.batch["read_off"])[-1])
    only_acgtn = bool(np.isin(np.unique(np.asarray(work.batch["read_bytes"])[:n_bases]), np.frombuffer(b"ACGTN", dtype=np.uint8)).all())
    prepared_packed = None
    if only_acgtn and os.environ.get("LTR_BENCH_PACKED", "1") != "0":
        packed_reads, keep_pk = pinned_copy(torch, dict(rb=abi.pack_reads_4bit(work.batch["read_bytes"], n_bases)))
        vb4, keep_vb4 = abi.make_viterbi_batch(dict(pinned_b, read_bytes=packed_reads["rb"]))
        prepared_packed = (vb4, pb, (keep_vb4, keep_pb, keep_pk))
    max_depth = int(os.environ.get("LTR_BENCH_DEPTH", "3"))
    outs = []
    for _ in range(max_depth):
        o, keep_o = pinned_copy(torch, dict(ll=np.zeros(n_ll), post=np.zeros(max(1, n_post)), tot=np.zeros(max(1, n_tot))))
        outs.append((o, keep_o))

    def e2e_run(depth, n_steps, prepared=prepared):
        inflight = collections.deque()
        last, acc = None, 0.0
        for i in range(n_steps + depth):
            if len(inflight) == depth or i >= n_steps:
                if not inflight:
                    break
                j, o = inflight.popleft()
                last = j.wait()
                acc += float(o["tot"][0])  # the host reads the step's result
                j.close()
            if i < n_steps:
                o = outs[i % depth][0]
                j = eng.submit_job(None, None, aln_params=work.aln_params, out_ll=o["ll"], out_post=o["post"][:n_post],
                                   out_totals=o["tot"][:n_tot], prepared=prepared)
                inflight.append((j, o))
        return last, acc
    e2e_by_depth = {}
    es = None
    for depth in range(1, max_depth + 1):
        e2e_run(depth, max(2, min(warmup, 3)))
        barrier(torch, world)
        t0 = time.perf_counter()
        es, _ = e2e_run(depth, steps)
        barrier(torch, world)
        e2e_by_depth[depth] = max_over_ranks(torch, world, (time.perf_counter() - t0) * 1e3 / steps)
    e2e_same = bool(np.array_equal(outs[0][0]["ll"], ll))  # the asynchronous path delivers the resident job's bits
    e2e_packed_by_depth, es4, packed_same = {}, None, None
    if prepared_packed is not None:
        eng.set_read_encoding(1)
        for o, _k in outs:
            o["ll"][...] = 0.0
        for depth in range(1, max_depth + 1):
            e2e_run(depth, max(2, min(warmup, 3)), prepared_packed)
            barrier(torch, world)
            t0 = time.perf_counter()
            es4, _ = e2e_run(depth, steps, prepared_packed)
            barrier(torch, world)
            e2e_packed_by_depth[depth] = max_over_ranks(torch, world, (time.perf_counter() - t0) * 1e3 / steps)
        eng.set_read_encoding(0)
        packed_same = bool(np.array_equal(outs[0][0]["ll"], ll))
    raw = None if args.no_raw else measure_from_raw_loci(args, config, n_loci, steps, torch, rank, world, local)
    in_flight = min(e2e_by_depth, key=e2e_by_depth.get)
    e2e_ms = e2e_by_depth[in_flight]
    # The headline end-to-end number is the plain form (one byte per base, nothing done to the caller's buffers); the 4-bit
    # stream -- packed by the caller outside the clock, as a caller that reads BAM records holds its reads -- is reported
    # beside it.
    e2e_encoding = "one byte per base"
    e2e_4bit_ms = min(e2e_packed_by_depth.values()) if (e2e_packed_by_depth and packed_same) else None
    if rank != 0:
        work.close()
        return None
    # roofline: cells the kernels actually evaluated (identical trimmed reads of a locus are aligned once)
    achieved = st.n_cells_computed / (vit_ms / steps) / 1e6
    line = {
        "metric": "loci_per_sec", "value": total_loci / (step_ms * 1e-3), "unit": "loci/s",
        "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": step_ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "gcups": total_cells / (vit_step_ms * 1e-3) / 1e9,
        "config": {"workload": CONFIG_NAME[config], "loci_per_gpu": n_loci,
                   "pairs_per_gpu": int(st.n_pairs), "cells_per_gpu": int(st.n_cells),
                   "pairs_aligned_per_gpu": int(st.n_pairs_computed), "cells_evaluated_per_gpu": int(st.n_cells_computed),
                   "pairs_banded_per_gpu": int(st.n_band_pairs), "pairs_band_uncertified_per_gpu": int(st.n_band_uncertified),
                   "pairs_band_second_round_per_gpu": int(st.n_band_retried),
                   "value_includes": "device plan (read de-duplication, band classes, task lists) + Viterbi kernels + "
                                     "fan-out + posteriors, every step; results start as NaN every step",
                   "plan_ms_per_step": plan_ms / steps, "viterbi_ms_per_step": vit_ms / steps,
                   "gcups_note": "gcups = reference-defined cells (every pooled read x haplotype) / Viterbi time; "
                                 "roofline.achieved = cells actually evaluated / Viterbi time",
                   "l2": "inputs (%.0f MB) + outputs larger than L2; no flush needed" % (work.input_bytes / 1e6),
                   "parallelism": "locus-sharded, no collective",
                   "rank_bound_to_cpus_of_its_gpu": NUMA_CPUS, "wall_ms_per_step": wall_step_ms,
                   "fallback_pairs": int(st.n_fallback), "ll_checksum": checksum},
        "e2e": {"value": total_loci / (e2e_ms * 1e-3), "unit": "loci/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(es.h2d_bytes), "d2h_bytes_per_step": int(es.d2h_bytes),
                "host_threads_per_gpu": 1, "api": "ltr_job_submit / ltr_job_wait",
                "jobs_in_flight": in_flight, "results_equal_resident_job": e2e_same,
                "read_encoding": e2e_encoding,
                "ms_per_step_by_jobs_in_flight": {str(k): v for k, v in e2e_by_depth.items()},
                "ms_per_step_by_jobs_in_flight_4bit_reads": {str(k): v for k, v in e2e_packed_by_depth.items()},
                "value_4bit_reads": (None if e2e_4bit_ms is None else total_loci / (e2e_4bit_ms * 1e-3)),
                "h2d_bytes_per_step_4bit_reads": (None if es4 is None else int(es4.h2d_bytes)),
                "note_4bit_reads": "ltr_ctx_set_read_encoding(1): reads as one 4-bit stream (BAM nibble codes), packed by the "
                                   "caller outside the clock; not the headline",
                "results_equal_resident_job_4bit_reads": packed_same},
        "e2e_from_flat_loci": raw,
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "fp64_issue", "achieved": achieved, "peak": peak_gcups, "unit": "GCUPS",
                     "frac": achieved / peak_gcups, "traffic": ncu_traffic("traffic"),
                     "traffic_source": ncu_traffic("source"), "traffic_kernel": ncu_traffic("kernel"),
                     "pipe_utilisation_ncu": ncu_traffic("pipe_utilisation_ncu"),
                     "peak_source": "ltr_fp64_issue_rate (DADD lane-ops/s measured in this run) / 17 FP64 ops per cell",
                     "fp64_lane_ops_per_s": fp64_rate,
                     "hbm_gbs_algorithmic": (work.input_bytes + 8.0 * st.n_pairs) / (vit_ms / steps) / 1e6},
    }
    if with_cpu_baseline:
        from oracle import pyoracle as po
        threads = os.cpu_count() or 1
        n_sample = args.cpu_sample_loci or min(n_loci, 256 * threads if config == 3 else 4 * threads)
        sb, sp = work.subset(n_sample)
        t0 = time.perf_counter()
        if po.ref_available():
            kind = "reference"
            ref_ll, sec, ref_post = po.ref_viterbi_batch(sb, work.aln_params, n_threads=threads, post=sp)
        else:
            kind = "port"
            ref_ll, _cells = po.viterbi_batch(sb, aln_params=work.aln_params, n_threads=threads)
            ref_post = None
            sec = time.perf_counter() - t0
        wall = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": n_sample / wall, "unit": "loci/s", "cores": threads, "kind": kind,
                                "gcups": sample_cells(work, n_sample) / sec / 1e9,
                                "sample": "first %d loci of the same workload, all host threads" % n_sample}
        line["parity_on_bench_sample"] = bench_sample_parity(work, n_sample, ref_ll, ref_post, ll, post)
    work.close()
    return line


def measure_from_raw_loci(args, config, n_loci, steps, torch, rank, world, local, devices=None):
    """The host front half inside the clock: raw loci (whole reads with CIGARs, flank blocks, candidate alleles; pageable
    host memory) -> ltr_genotyper_run = pooling + trimming + flattening on the host threads, asynchronous GPU jobs
    (Viterbi, posteriors, removal of uncalled alleles), call extraction -> GT / Q / PQ / GL / PL per sample."""
    from longtr_b200 import Genotyper, workloads
    threads = max(1, (os.cpu_count() or 1) // max(1, world))
    work = workloads.generate_loci(config, n_loci, first_locus=rank * n_loci)
    gen = Genotyper(devices=tuple(devices) if devices else (local,), host_threads=threads)
    n_steps = max(1, min(steps, 3))
    calls = gen.run_struct(work.struct, work.aln_params)   # warm-up: pinned buffers, memory pool
    gen.free(calls)
    barrier(torch, world)
    t0 = time.perf_counter()
    prep = wait = post = submit = 0.0
    n_ok = n_chunks = 0
    for _ in range(n_steps):
        calls = gen.run_struct(work.struct, work.aln_params)
        c = calls.contents
        submit += c.submit_ms
        n_chunks = c.n_chunks
        prep += c.prep_ms
        wait += c.gpu_wait_ms
        post += c.post_ms
        n_ok = int(np.sum(np.ctypeslib.as_array(c.status, (n_loci,)) == 0))
        gen.free(calls)
    barrier(torch, world)
    ms = max_over_ranks(torch, world, (time.perf_counter() - t0) * 1e3 / n_steps)
    total_loci = sum_over_ranks(torch, world, float(n_loci))
    out = {"value": total_loci / (ms * 1e-3), "unit": "loci/s", "ms_per_step": ms, "steps": n_steps,
           "api": "ltr_genotyper_run (raw loci in pageable host memory -> calls)", "host_threads_per_gpu": threads,
           "host_prepare_ms_per_step": prep / n_steps, "host_wait_for_gpu_ms_per_step": wait / n_steps,
           "host_extract_calls_ms_per_step": post / n_steps, "host_submit_ms_per_step": submit / n_steps,
           "jobs_per_step": int(n_chunks), "loci_genotyped": n_ok,
           "input_bytes_per_step": work.input_bytes, "reads_per_step": int(work.n_reads)}
    gen.close()
    work.close()
    return out


N2_PEAK_GCUPS = 148 * 2 * 1.965e9 / 17.0 * 1024 / 1e9
CLUSTER_THRESHOLDS = [20, 50, 80, 100, 150, 200, 300, 400, 500, 600, 700]  # HaplotypeGenerator.cpp:403


def measure_cluster(args, eng, n_sets, steps, warmup, with_cpu_baseline):
    """extra.n2: greedy clustering of the skipped sequences of n_sets (locus, sample) pairs at every threshold of the
    reference's ladder in one call (ltr_cluster_greedy), host buffers in, assignments out.  The reference walks the ladder
    until greedy_clustering succeeds; the CPU arm does exactly that on a bounded sample."""
    from longtr_b200.workloads import generate_cluster_sets
    seq_bytes, seq_off, begin = generate_cluster_sets(n_sets)
    nT = len(CLUSTER_THRESHOLDS)
    items, set_begin, set_T = [], [0], []
    for k in range(n_sets):
        ids = np.arange(begin[k], begin[k + 1], dtype=np.uint32)
        for T in CLUSTER_THRESHOLDS:
            items.append(ids)
            set_begin.append(set_begin[-1] + len(ids))
            set_T.append(T)
    items = np.concatenate(items)
    ms, kms = [], []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        cent, ncent, ok, st = eng.cluster_greedy(seq_bytes, seq_off, set_begin, items, set_T)
        if it >= warmup:
            ms.append((time.perf_counter() - t0) * 1e3)
            kms.append(st.kernel_ms)
    ok = ok.reshape(n_sets, nT)
    first_ok = np.where(ok.any(axis=1), ok.argmax(axis=1), -1)
    # comparisons the rounds made: every item behind each centroid of its set
    lens = np.diff(seq_off.astype(np.int64))
    line = {"metric": "cluster_sets_per_sec", "value": n_sets / (np.mean(ms) / 1e3), "unit": "sets/s",
            "ms_per_step": float(np.mean(ms)), "kernel_ms_per_step": float(np.mean(kms)), "steps": steps, "warmup": warmup,
            "api": "ltr_cluster_greedy (host buffers in, assignments out; all 11 thresholds of a set in one call)",
            "gcups": float(st.n_cells) / (np.mean(kms) * 1e-3) / 1e9, "comparisons_per_step": int(st.n_pairs),
            "roofline": {"bound": "alu_issue", "achieved": float(st.n_cells) / (np.mean(kms) * 1e-3) / 1e9,
                         "peak": N2_PEAK_GCUPS, "unit": "GCUPS", "frac": float(st.n_cells) / (np.mean(kms) * 1e-3) / 1e9 / N2_PEAK_GCUPS,
                         "traffic": None,
                         "peak_source": "148 SMs x 2 ALU-pipe warp instructions per clock x 1.965 GHz / 17 word operations of "
                                        "Myers' recurrence per 32-row word x 1024 cells per warp step; achieved counts the "
                                        "reference-defined n*m cells of every comparison over the device time of the call "
                                        "(15 rounds, 61 launches)"},
            "gpu_launches": int(st.n_launches),
            "config": {"workload": "N2: skipped sequences of (locus, sample) pairs, config-4-like VNTR alleles (500-1000 bp, "
                                   "2-4 alleles, ~40 distinct noisy copies), thresholds 20..700",
                       "sets": n_sets, "sequences": int(len(lens)), "set_thresholds": int(n_sets * nT),
                       "mean_len": float(lens.mean()), "sets_clustered_at_first_threshold_index": np.bincount(
                           first_ok[first_ok >= 0], minlength=nT).tolist()}}
    if with_cpu_baseline:
        from oracle import pyoracle as po
        import concurrent.futures as cf
        threads = os.cpu_count() or 1
        n_sample = min(n_sets, threads * 2)
        which = "ref" if po.ref_available() else "oracle"

        def ladder(k):
            ids = np.arange(begin[k], begin[k + 1], dtype=np.uint32)
            for ti, T in enumerate(CLUSTER_THRESHOLDS):
                okk, c, n = po.greedy_cluster(seq_bytes, seq_off, ids, T, which)
                if okk:
                    return ti, c
            return -1, None
        t0 = time.perf_counter()
        with cf.ThreadPoolExecutor(threads) as ex:
            res = list(ex.map(ladder, range(n_sample)))
        sec = time.perf_counter() - t0
        same, n_diff = True, 0
        for k, (ti, c) in enumerate(res):
            ok_k = (ti == int(first_ok[k]))
            if ok_k and ti >= 0:
                b0 = set_begin[k * nT + ti]
                ok_k = bool(np.array_equal(c, cent[b0:b0 + len(c)]))
            same &= ok_k
            n_diff += (not ok_k)
        line["cpu_baseline"] = {"value": n_sample / sec, "unit": "sets/s", "cores": threads,
                                "kind": "reference" if which == "ref" else "port",
                                "sample": "first %d sets, threshold ladder until greedy_clustering succeeds "
                                          "(HaplotypeGenerator.cpp:403-409), all host threads" % n_sample}
        line["parity_on_bench_sample"] = {"sets": n_sample, "first_threshold_and_assignments_equal": bool(same),
                                          "sets_that_differ": int(n_diff)}
    return line


def measure_regions(args, n_loci, steps, warmup):
    """extra.n3: BAM files -> calls through ltr_regions_run (BGZF / BAM / BAI reader, read filters, trimming, candidate
    alleles on host threads; alignment, posteriors, call extraction through ltr_genotyper_run).  The BAM file is written
    here from the raw loci of the config-3 generator (tests/bam_writer.py); regions whose reads are not explained by exact
    candidates get consensus alleles from the assembly branch (clustering + partial-order consensus on the host threads,
    DESIGN.md section 6c) and are genotyped like the others."""
    import tempfile
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests"))
    import bam_writer as bw
    from longtr_b200 import abi
    from longtr_b200.locus_batch import Genotyper
    t0 = time.perf_counter()
    world = bw.synthetic_world(n_loci, config=3, n_samples=1)
    d = tempfile.mkdtemp(prefix="ltr_n3_")
    paths = bw.write_world(world, d)
    gen_s = time.perf_counter() - t0
    bams = [abi.BamFile(p) for p in paths]
    t0 = time.perf_counter()
    for b in bams:
        b.build_index()
    index_ms = (time.perf_counter() - t0) * 1e3
    g = Genotyper(devices=(0,), host_threads=0, chunk_loci=0)
    ms, wall, out = [], [], None
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        out = g.run_regions(bams, "chrS", world["regions"], world["chrom_seq"], 0)
        if it >= warmup:
            ms.append(out["call_ms"])   # the C call; the harness' conversion of the result into Python objects is outside
            wall.append((time.perf_counter() - t0) * 1e3)
    # the same with the VCF record of every region composed as well (per-read allele assignment: LL matrices downloaded)
    motifs = [world["chrom_seq"][s0:s0 + per] for s0, _e, per in world["regions"]]
    ms_rec, n_rec = [], 0
    for it in range(1 + max(1, steps)):
        t0 = time.perf_counter()
        o2 = g.run_regions(bams, "chrS", world["regions"], world["chrom_seq"], 0, motifs=motifs)
        if it >= 1:
            ms_rec.append(o2["call_ms"])
        n_rec = sum(1 for x in o2["records"] if x)
    g.close()
    status = np.array(out["status"])
    bam_bytes = sum(os.path.getsize(p) for p in paths)
    for p in paths:
        os.remove(p)
    os.rmdir(d)
    t = out["calls"]["timing"] if out["calls"] else {}
    return {"metric": "regions_per_sec", "value": n_loci / (np.mean(ms) / 1e3), "unit": "regions/s",
            "ms_per_step": float(np.mean(ms)), "steps": steps, "warmup": warmup,
            "api": "ltr_regions_run (BAM file + regions + reference sequence in, calls out)",
            "host_threads_per_gpu": os.cpu_count(),
            "with_vcf_records": {"value": n_loci / (np.mean(ms_rec) / 1e3), "unit": "regions/s", "records": int(n_rec),
                                 "note": "ltr_regions_run with opts.vcf_records (LL download, per-read alleles, record text)"},
            "wall_ms_per_step_incl_python_decoding": float(np.mean(wall)),
            "genotyper_ms": {k: float(v) for k, v in t.items()},
            "host_ms": {k: float(v) for k, v in out.get("host_ms", {}).items()},
            "host_ms_with_vcf_records": {k: float(v) for k, v in o2.get("host_ms", {}).items()},
            "config": {"workload": "N3: %d config-3 loci as one coordinate-sorted BAM file (30 spanning reads per region, "
                                   "1.5 kb each), one sample" % n_loci, "regions": n_loci,
                       "regions_genotyped": int((status == 0).sum()), "regions_assembled": int(out["n_assembled"]),
                       "consensus_alleles": int(sum(map(sum, out["inexact"]))),
                       "regions_other": int((status != 0).sum()), "bam_bytes": int(bam_bytes),
                       "reads": int(sum(len(r) for r in world["records"])), "index_build_ms": index_ms,
                       "bam_written_in_s": gen_s}}


def measure_em(args, eng, n_loci, steps, warmup, with_cpu_baseline):
    """extra.em: length-based EM of the stutter model (ltr_em_stutter_train, one warp per locus) on seeded loci of 1-3 samples
    with 3-40 reads each, host buffers in, parameters out; the reference's EMStutterGenotyper (compiled in place) on one host
    thread over a bounded sample beside it, and the parameters of that sample compared."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests"))
    import em_cases
    from longtr_b200 import abi
    base = [em_cases.em_locus(100000 + k) for k in range(min(n_loci, 2000))]
    loci = [base[k % len(base)] for k in range(n_loci)]
    packed = abi.em_pack(loci)
    ms, got = [], None
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        got = abi.em_stutter_train(eng.ctx, packed)
        if it >= warmup:
            ms.append((time.perf_counter() - t0) * 1e3)
    n_reads = sum(sum(L["reads_per_sample"]) for L in loci)
    line = {"metric": "loci_per_sec", "value": n_loci / (np.mean(ms) / 1e3), "unit": "loci/s", "ms_per_step": float(np.mean(ms)),
            "steps": steps, "warmup": warmup, "api": "ltr_em_stutter_train (host buffers in, parameters out)", "config": {"workload": "N4: stutter-model EM, %d loci, %d reads" % (n_loci, n_reads)},
            "iterations_mean": float(np.mean(got["n_iter"])), "trained_fraction": float(np.mean(got["trained"]))}
    if with_cpu_baseline:
        from oracle import pyoracle as po
        if po.ref_em_available():
            n_s = min(400, len(base))
            t0 = time.perf_counter()
            ref = [po.ref_em_train(L["reads_per_sample"], L["bp_diff"], L["log_p1"], L["log_p2"], L["motif_len"], L["haploid"])
                   for L in base[:n_s]]
            sec = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": n_s / sec, "unit": "loci/s", "cores": 1, "kind": "reference",
                                    "sample": "first %d loci, EMStutterGenotyper compiled in place, one thread" % n_s}
            err = max(float(np.max(np.abs(got["params"][k] - ref[k]["params"]) / np.maximum(np.abs(ref[k]["params"]), 1e-300)))
                      for k in range(n_s))
            line["parity_on_bench_sample"] = {"loci": n_s, "max_rel_err_params": err,
                                              "same_iterations": bool(all(got["n_iter"][k] == ref[k]["n_iter"] for k in range(n_s))),
                                              "same_trained": bool(all(bool(got["trained"][k]) == ref[k]["trained"] for k in range(n_s)))}
    return line


def compact(line):
    """Sub-line of another configuration inside the default run (extra.c4 / extra.c5)."""
    if line is None:
        return None
    keep = ("metric", "value", "unit", "ms_per_step", "gcups", "steps", "warmup", "e2e", "e2e_from_flat_loci", "roofline",
            "cpu_baseline",
            "parity_on_bench_sample", "gpu_launches", "clocks")
    out = {k: line[k] for k in keep if k in line}
    out["config"] = {k: line["config"][k] for k in ("workload", "loci_per_gpu", "pairs_per_gpu", "pairs_aligned_per_gpu",
                                                     "cells_evaluated_per_gpu", "pairs_banded_per_gpu",
                                                     "pairs_band_uncertified_per_gpu", "pairs_band_second_round_per_gpu",
                                                     "cell_equivalents_per_gpu")
                     if k in line["config"]}
    return out


def main():
    args = parse_args()
    if args.impl == "reference":
        if args.config == 5:
            run_reference_stutter(args)
        else:
            run_reference(args)
        return
    torch, rank, world, local = dist_setup(args.gpus)
    from longtr_b200 import Engine
    eng = Engine(local)
    if args.one_process_devices:
        # the product-level multi-GPU form: one process, one context per device, chunks round robin, results in input order
        nd = args.one_process_devices
        line = {"metric": "loci_per_sec", "unit": "loci/s", "n_gpus": nd, "one_process": True}
        for k in (1, nd):
            r = measure_from_raw_loci(args, args.config, (args.loci or CONFIG_LOCI[args.config]) * k, args.steps, torch, rank,
                                      world, local, devices=range(k))
            line["devices_%d" % k] = r
        line["value"] = line["devices_%d" % nd]["value"]
        line["speedup_over_one_device"] = line["value"] / line["devices_1"]["value"]
    elif args.n2:
        line = measure_cluster(args, eng, args.loci or 512, args.steps, args.warmup, not args.no_cpu_baseline)
    elif args.n3:
        line = measure_regions(args, args.loci or 1500, args.steps, args.warmup)
    elif args.em:
        line = measure_em(args, eng, args.loci or 20000, args.steps, args.warmup, not args.no_cpu_baseline)
    elif args.config == 5:
        line = run_stutter(args, torch, rank, world, local, eng, args.loci or CONFIG_LOCI[5], args.steps, args.warmup,
                           world == 1 and not args.no_cpu_baseline)
    else:
        # roofline denominator: sustained FP64-pipe issue rate measured on this GPU, right now
        fp64_rate = max(eng.fp64_issue_rate(0)[0] for _ in range(2))
        peak_gcups = fp64_rate / FP64_OPS_PER_CELL / 1e9
        line = measure_long(args, args.config, args.loci or CONFIG_LOCI[args.config], args.steps, args.warmup, torch, rank,
                            world, local, eng, peak_gcups, fp64_rate, world == 1 and not args.no_cpu_baseline)
        # The default single-GPU run also times the other two synthetic configurations (shorter runs, bounded CPU
        # samples) so that their numbers are driver-run too: extra.c4 (VNTRs, ONT-like) and extra.c5 (homopolymer path).
        if args.config == 3 and world == 1 and not args.loci and not args.no_extra:
            extra = {}
            try:
                extra["c4"] = compact(measure_long(args, 4, CONFIG_LOCI[4], 2, 3, torch, rank, world, local, eng, peak_gcups,
                                                   fp64_rate, not args.no_cpu_baseline))
                extra["c5"] = compact(run_stutter(args, torch, rank, world, local, eng, CONFIG_LOCI[5], 2, 3,
                                                  not args.no_cpu_baseline))
                extra["n2"] = measure_cluster(args, eng, 512, 3, 3, not args.no_cpu_baseline)
                extra["n3"] = measure_regions(args, 1500, 2, 1)
                extra["em"] = measure_em(args, eng, 20000, 2, 1, not args.no_cpu_baseline)
            except Exception as e:  # the headline line must not be lost to a sub-line
                extra["error"] = repr(e)
            line["extra"] = extra
    if rank == 0 and line is not None:
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
