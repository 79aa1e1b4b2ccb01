#!/usr/bin/env python
"""Benchmark of the quasi-mapping hot path (BASELINE.json: paired-end 2x100 bp read pairs/s; SA-lookup kernel
HBM GB/s vs peak).

    python bench.py --gpus N --steps K --warmup W            # this engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU `quasimap` on the host cores

Workload (configs[1]): GENCODE-like ~200k-transcript synthetic index (tools/synth.cpp, seed 12345, 37,000 genes),
synthetic 2x100 bp pairs (seed 54321: 1% substitutions, 0.03% ins/del, 0.1% N), default `quasimap` flags (no -s).
A step = one chunk ("batch") of --batch pairs per GPU through rapmap_cuda_map_batch; 40 default steps of 2^20
pairs (4 distinct batches cycled) = 4x the 10M-pair configuration, so that the timed region is long enough to sample clocks.  Index replicated per GPU (one NCCL broadcast of the packed image), read
ranges sharded by rank, no data-path collective: weak scaling.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import numpy as np  # noqa: E402

METRIC = "paired-end 2x100bp read pairs/s (quasimap hot path)"
UNIT = "pairs/s"
READ_LEN = 100
TX_SEED, READ_SEED = 12345, 54321
CACHE = os.environ.get("RAPMAP_B200_CACHE", "/tmp/rapmap_b200_cache")


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


# The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on the first
# collective), so file descriptor 1 is pointed at stderr for the whole run and the line goes to the saved descriptor.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict) -> None:
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def index_dir(genes: int) -> str:
    return os.path.join(CACHE, f"bench_g{genes}_s{TX_SEED}", "idx")


def ensure_index(genes: int, use_gpu: bool) -> str:
    """Reference-format index of the synthetic transcriptome (tools/build_index.py), cached per box under CACHE."""
    d = index_dir(genes)
    if os.path.exists(os.path.join(d, "header.json")):
        return d + "/"
    from build_index import build_synth_index

    tmp = d + ".tmp%d" % os.getpid()
    t0 = time.time()
    build_synth_index(tmp, TX_SEED, genes, 0, device="cuda" if use_gpu else "cpu", verbose=False)
    os.makedirs(os.path.dirname(d), exist_ok=True)
    try:
        os.rename(tmp, d)
    except OSError:
        pass
    log(f"built index for {genes} genes in {time.time() - t0:.1f}s -> {d}")
    return d + "/"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def alg_bytes_per_pair(ops: dict, pairs: int) -> float:
    """Algorithmic bytes of the SA-lookup kernel per pair (payload granularity):
    2*L read bases + 16 B per k-mer lookup + 4 B per SA probe + 1 B per text character compared."""
    return 2 * READ_LEN + (16.0 * ops["hashFind"] + 4.0 * ops["saProbes"] + 1.0 * ops["textCmp"]) / pairs


def run_reference(args) -> None:
    """The reference's own multithreaded CPU `quasimap` (oracle/_ref/rapmap_ref, unmodified sources) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from helpers import REF_BIN, SynthTxome

    cores = os.cpu_count() or 1
    line = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic"}
    if not os.path.exists(REF_BIN):
        line["unavailable"] = "oracle/_ref/rapmap_ref not built (needs /root/reference at build time)"
        emit(line)
        return
    import torch

    idx = ensure_index(args.genes, torch.cuda.is_available())
    sample = args.ref_pairs_per_step
    total = sample * args.steps
    tx = SynthTxome(TX_SEED, args.genes)
    d = os.path.join(CACHE, "ref_reads")
    os.makedirs(d, exist_ok=True)
    f1, f2 = os.path.join(d, f"r1_{total}.fastq"), os.path.join(d, f"r2_{total}.fastq")
    if not os.path.exists(f2):
        qual = b"I" * READ_LEN
        with open(f1 + ".tmp", "wb") as g1, open(f2 + ".tmp", "wb") as g2:
            for first in range(0, total, 500000):  # chunked: bounded host memory whatever --steps says
                cnt = min(500000, total - first)
                s1, s2 = tx.reads(cnt, rseed=READ_SEED, first=first, read_len=READ_LEN)
                for g, arr, m in ((g1, s1, 1), (g2, s2, 2)):
                    g.write(b"".join(b"@r%d/%d\n%s\n+\n%s\n" % (first + i, m, arr[i].tobytes(), qual) for i in range(cnt)))
        os.rename(f1 + ".tmp", f1)
        os.rename(f2 + ".tmp", f2)
    flags = ["-s"] if args.selaln else []

    def once(n_pairs_files):
        t0 = time.time()
        p = subprocess.run([REF_BIN, "quasimap", "-i", idx, "-1", n_pairs_files[0], "-2", n_pairs_files[1], "-t", str(cores), "-n"] + flags,
                           capture_output=True, text=True)
        wall = time.time() - t0
        m = re.findall(r"Elapsed time: ([0-9.eE+-]+)s", p.stdout + p.stderr)
        return (float(m[-1]) if m else wall), wall

    if args.warmup > 0:  # one untimed pass warms the page cache (the CLI is one-shot: warm-up steps cannot be separated)
        w1, w2 = os.path.join(d, "w1.fastq"), os.path.join(d, "w2.fastq")
        for src, dst in ((f1, w1), (f2, w2)):
            with open(src, "rb") as f, open(dst, "wb") as g:
                g.write(b"".join(f.readline() for _ in range(4 * 2000)))
        once((w1, w2))
    mapping_s, wall = once((f1, f2))
    value = total / mapping_s
    line.update({
        "value": value, "ms_per_step": 1e3 * mapping_s / args.steps,
        "config": {"workload": f"GENCODE-like {args.genes}-gene synthetic index, 2x100bp synthetic pairs, quasimap{' -s' if args.selaln else ''} -n -t {cores}",
                   "pairs_per_step": sample, "timed": "reference ScopedTimer 'Elapsed time' around mapReads (index load excluded)", "wall_s_incl_index_load": wall},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": f"{total} pairs (first pairs of the benchmark read stream)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1 << 20, help="pairs per step per GPU")
    ap.add_argument("--genes", type=int, default=37000, help="synthetic genes (37000 -> ~203k transcripts)")
    ap.add_argument("--selaln", action="store_true", help="quasimap -s (configs[2])")
    ap.add_argument("--distinct", type=int, default=4, help="distinct read batches cycled through the steps")
    ap.add_argument("--ref-pairs-per-step", type=int, default=100000, help="pairs per step of the reference arm (bounded sample: 40 steps = 4M pairs, ~5 s of 16-thread CPU mapping)")
    ap.add_argument("--oracle-sample", type=int, default=20000, help="pairs checked against / counted by the CPU oracle at N=1")
    ap.add_argument("--cpu-baseline-pairs", type=int, default=1000000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-mappers", type=int, default=4, help="host threads (one mapper / CUDA stream each) of the end-to-end measurement")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    import rapmap_b200 as rb
    from helpers import REF_BIN, OracleMapper, SynthTxome

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()

    # ---- index: rank 0 builds (cached) + loads from the reference-format files; other ranks receive the packed image over NCCL
    if rank == 0:
        idx_dir = ensure_index(args.genes, True)
    barrier()
    idx_dir = index_dir(args.genes) + "/"
    t0 = time.time()
    from rapmap_b200.sharding import replicate_index

    index = rb.Index(idx_dir, local) if rank == 0 else None
    index, _image_keepalive = replicate_index(index, rank, local)
    log(f"rank {rank}: index ready in {time.time() - t0:.1f}s ({index.device_bytes / 2**30:.2f} GiB in HBM, {index.num_transcripts} transcripts)")

    # ---- reads: rank-sharded contiguous ranges of the counter-based stream; pinned host copies + device copies
    B = args.batch
    opts = rb.default_opts(sel_aln=args.selaln)
    mapper = rb.Mapper(index, opts, max_batch=B, max_read_len=READ_LEN)
    tx = SynthTxome(TX_SEED, args.genes)
    nd = max(1, min(args.distinct, args.steps + args.warmup))
    host, devb = [], []
    t0 = time.time()
    for b in range(nd):
        h1 = torch.empty((B, READ_LEN), dtype=torch.uint8).pin_memory()
        h2 = torch.empty((B, READ_LEN), dtype=torch.uint8).pin_memory()
        first = (rank * nd + b) * B
        tx.reads(B, rseed=READ_SEED, first=first, read_len=READ_LEN, out1=h1.numpy(), out2=h2.numpy())
        host.append((h1, h2))
        devb.append((h1.cuda(), h2.cuda()))
    log(f"rank {rank}: generated {nd} x {B} pairs in {time.time() - t0:.1f}s")
    cap = B * 8
    d_hits = torch.empty(cap * 28, dtype=torch.uint8, device="cuda")
    d_off = torch.empty(B + 1, dtype=torch.int64, device="cuda")
    h_hits = torch.empty(cap * 28, dtype=torch.uint8).pin_memory()
    h_off = torch.empty(B + 1, dtype=torch.int64).pin_memory()
    stream = torch.cuda.ExternalStream(mapper.stream_ptr, device=torch.device("cuda", local))

    def step_resident(i):
        a, b = devb[i % nd]
        return mapper.map_batch(a, b, n=B, fixed_len=READ_LEN, location=rb.LOC_DEVICE, hits_out=d_hits, offsets_out=d_off, out_location=rb.LOC_DEVICE, capacity=cap)

    def timed(fn, steps, warmup, sample_clocks=False):
        for i in range(warmup):
            fn(i)
        torch.cuda.synchronize()
        barrier()
        sampler = ClockSampler(local) if sample_clocks else None
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        stages = {"pack": 0.0, "sa": 0.0, "map": 0.0, "merge": 0.0, "selaln": 0.0, "h2d": 0.0, "d2h": 0.0, "launch": 0, "hits": 0, "retries": 0}
        for i in range(steps):
            r = fn(warmup + i)
            t = mapper.timing()
            stages["pack"] += t.ms_pack_reads; stages["sa"] += t.ms_sa_collect; stages["map"] += t.ms_hits_to_mappings; stages["merge"] += t.ms_merge; stages["selaln"] += t.ms_sel_aln
            stages["h2d"] += t.ms_h2d; stages["d2h"] += t.ms_d2h; stages["launch"] += t.launches; stages["hits"] += r.num_hits; stages["retries"] += t.retries
        e1.record(stream)
        torch.cuda.synchronize()
        barrier()
        clocks = sampler.stop() if sampler else None
        ms = e0.elapsed_time(e1)
        if world > 1:
            tm = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            ms = float(tm.item())
        return ms, stages, clocks

    ms_res, st_res, clocks = timed(step_resident, args.steps, args.warmup, sample_clocks=True)
    log(f"rank {rank}: resident leg {ms_res / args.steps:.2f} ms per step")

    # ---- end to end: HOST buffers in, HOST buffers out, through rapmap_cuda_map_batch.  Like the reference's worker threads
    # (one SACollector per thread), --e2e-mappers host threads each own a mapper (= one CUDA stream) and take chunks
    # round-robin, so one chunk's PCIe copies overlap another chunk's kernels.  Every call is synchronous for its caller.
    nm = max(1, args.e2e_mappers)
    lanes = [(mapper, h_hits, h_off)]
    for _ in range(nm - 1):
        lanes.append((rb.Mapper(index, opts, max_batch=B, max_read_len=READ_LEN), torch.empty(cap * 28, dtype=torch.uint8).pin_memory(),
                      torch.empty(B + 1, dtype=torch.int64).pin_memory()))

    def e2e_worker(t, first, count, acc):
        torch.cuda.set_device(local)
        mp, hh, ho = lanes[t]
        hits = 0
        for i in range(first + t, first + count, nm):
            a, b = host[i % nd]
            r = mp.map_batch(a.numpy(), b.numpy(), n=B, fixed_len=READ_LEN, location=rb.LOC_HOST, hits_out=hh, offsets_out=ho, out_location=rb.LOC_HOST, capacity=cap)
            hits += r.num_hits
        acc[t] = hits

    def run_e2e(first, count):
        acc = [0] * nm
        ths = [threading.Thread(target=e2e_worker, args=(t, first, count, acc)) for t in range(nm)]
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        return sum(acc)

    w_e2e = max(3, args.warmup)
    run_e2e(0, w_e2e * nm)
    torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_hits = run_e2e(w_e2e * nm, args.steps)
    e1.record()
    torch.cuda.synchronize()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        tm = torch.tensor([ms_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms_e2e = float(tm.item())
    st_e2e = {"hits": e2e_hits}
    log(f"rank {rank}: end-to-end leg {ms_e2e / args.steps:.2f} ms per step ({nm} host threads)")
    total_pairs = B * args.steps * world
    value = total_pairs / (ms_res / 1e3)
    e2e_value = total_pairs / (ms_e2e / 1e3)

    if rank == 0:
        peak, peak_src = measured_peaks()
        # ---- oracle leg: parity check + operation counts on a bounded sample (N == 1 only)
        ops_per_pair, parity = None, None
        alg = None
        if world == 1 and args.oracle_sample > 0:
            ns = min(args.oracle_sample, B)
            a, b = host[0][0].numpy()[:ns].copy(), host[0][1].numpy()[:ns].copy()
            om = OracleMapper(idx_dir, opts)
            t0 = time.time()
            ref = om.map(a, b, READ_LEN)
            port_s = time.time() - t0
            got = mapper.map_batch(a, b, n=ns, fixed_len=READ_LEN)
            parity = bool(got.num_hits == ref.num_hits and np.array_equal(got.hits, ref.hits) and np.array_equal(got.pair_offsets, ref.pair_offsets)
                          and np.array_equal(got.counters, ref.counters))
            ops = om.op_counts()
            ops_per_pair = {k: v / ns for k, v in ops.items()}
            alg = alg_bytes_per_pair(ops, ns)
            log(f"oracle sample: {ns} pairs in {port_s:.1f}s, parity={parity}, ops/pair={ops_per_pair}")
            if not parity:
                raise SystemExit("PARITY FAILURE against the CPU oracle on the benchmark sample; refusing to report a number")
        if alg is None:
            alg = 2 * READ_LEN + 16.0 * 96.4 + 4.0 * 17.7 + 654.0  # counted on this workload (no -s)
        sa_ms = st_res["sa"] / args.steps
        achieved = alg * B / (sa_ms / 1e3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "sa_collect_traffic_selaln.json" if args.selaln else "sa_collect_traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        # ---- cpu baseline: the reference binary on the host cores (bounded sample)
        cpu = None
        if world == 1 and not args.no_cpu_baseline and os.path.exists(REF_BIN):
            ns = args.cpu_baseline_pairs
            cores = os.cpu_count() or 1
            d = os.path.join(CACHE, "cpu_reads")
            os.makedirs(d, exist_ok=True)
            f1, f2 = os.path.join(d, f"r1_{ns}.fastq"), os.path.join(d, f"r2_{ns}.fastq")
            if not os.path.exists(f2):
                s1, s2 = tx.reads(ns, rseed=READ_SEED, first=0, read_len=READ_LEN)
                qual = b"I" * READ_LEN
                for path, arr, m in ((f1, s1, 1), (f2, s2, 2)):
                    with open(path, "wb") as f:
                        f.write(b"".join(b"@r%d/%d\n%s\n+\n%s\n" % (i, m, arr[i].tobytes(), qual) for i in range(ns)))
            p = subprocess.run([REF_BIN, "quasimap", "-i", idx_dir, "-1", f1, "-2", f2, "-t", str(cores), "-n"] + (["-s"] if args.selaln else []),
                               capture_output=True, text=True)
            m = re.findall(r"Elapsed time: ([0-9.eE+-]+)s", p.stdout + p.stderr)
            if m:
                cpu = {"value": ns / float(m[-1]), "unit": UNIT, "cores": cores, "kind": "reference",
                       "sample": f"first {ns} pairs of the benchmark stream, rapmap_ref quasimap -n -t {cores}, reference ScopedTimer (index load excluded)"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {"workload": f"configs[1]: GENCODE-like {args.genes}-gene (~{index.num_transcripts} txp) synthetic index, 2x100bp synthetic pairs, quasimap{' -s' if args.selaln else ''} default flags",
                       "pairs_per_step_per_gpu": B, "distinct_batches": nd, "l2_policy": f"inputs larger than L2 ({2 * B * READ_LEN / 2**20:.0f} MiB of read bases per step, {index.device_bytes / 2**30:.1f} GiB index)",
                       "index": "replicated per GPU (one NCCL broadcast of the packed image)", "sharding": "contiguous read ranges per rank, no data-path collective"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * B * READ_LEN, "d2h_bytes_per_step": int(st_e2e["hits"] / args.steps * 28 + (B + 1) * 8),
                    "ms_per_step": ms_e2e / args.steps, "api": "rapmap_cuda_map_batch with pinned HOST buffers", "host_threads": nm,
                    "note": "one mapper (CUDA stream) per host thread, chunks round-robin; each call synchronous for its caller"},
            "gpu_launches": int(st_res["launch"]),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "sa_collect_lane_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "alg_bytes_per_pair": alg, "kernel_ms_per_launch": sa_ms,
                         "stage_ms_per_step": {k: st_res[k] / args.steps for k in ("pack", "sa", "map", "merge", "selaln", "h2d", "d2h")},
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel at this batch size (profiles/)" if traffic else None},
            "cpu_baseline": cpu,
            "hits_per_pair": st_res["hits"] / (B * args.steps),
            "ops_per_pair": ops_per_pair,
            "parity_checked_vs_oracle": parity,
        }
        emit(line)
        log(f"rank 0: {value / 1e6:.1f} M pairs/s resident, {e2e_value / 1e6:.1f} M pairs/s end to end on {world} GPU(s)")
    barrier()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
