#!/usr/bin/env python
"""Benchmark of the quasi-mapping hot path (BASELINE.json: paired-end 2x100 bp read pairs/s; SA-lookup kernel
HBM GB/s vs peak).

    python bench.py --gpus N --steps K --warmup W            # this engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU `quasimap` on the host cores

Workload: GENCODE-like ~203k-transcript synthetic index (tools/synth.cpp, seed 12345, 37,000 genes), synthetic 2x100 bp
pairs (seed 54321: 1% substitutions, 0.03% ins/del, 0.1% N).  A STEP is one pass over the 10M-read configuration of
BASELINE.json configs[1]: --chunks (10) chunks of --batch (2^20) pairs per GPU, each chunk one rapmap_cuda_map_batch call
(4 distinct chunks cycled; every chunk's bases are larger than L2).  The ONE JSON line carries

  * the headline: configs[1], default `quasimap` flags (no -s): `value` with reads resident in HBM, `e2e` through the C-ABI
    with pinned HOST buffers (copies inside the timed region), the SA-lookup kernel's roofline, the reference CPU baseline;
  * `legs.selaln`       configs[2] (and, under torchrun, configs[4]): the same index and reads with `quasimap -s`;
  * `legs.perfect_hash` configs[3]: the same transcriptome indexed with -p (BooPHF + FrugalBooMap; lookups served by the dense
    table the engine derives from FrugalBooMap::find when it loads the index; `walk`: the same leg with the BooPHF walked per lookup).

Every leg is parity-checked before its number is printed (CPU oracle on a bounded sample; -p against the dense results).
Index replicated per GPU (one NCCL broadcast of the packed image), read ranges sharded by rank, no data-path collective:
weak scaling.
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


def workload_name(genes: int) -> str:
    """The same string in both arms (this engine and --impl reference)."""
    return f"configs[1]: GENCODE-like {genes}-gene (~203k-transcript at 37000 genes) synthetic index, synthetic 2x100bp read pairs, quasimap default flags (no -s)"


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


def index_dir(genes: int, perfect: bool = False) -> str:
    return os.path.join(CACHE, f"bench_g{genes}_s{TX_SEED}", "idx_p" if perfect else "idx")


def ensure_index(genes: int, use_gpu: bool, perfect: bool = False) -> str:
    """Reference-format index of the synthetic transcriptome (tools/build_index.py: dense flavour and the -p flavour of the same
    suffix array, written in one go), cached per box under CACHE."""
    d, dp = index_dir(genes), index_dir(genes, True)
    want = dp if perfect else d
    if os.path.exists(os.path.join(want, "header.json")):
        return want + "/"
    from build_index import build_synth_index

    tmp, tmpp = d + ".tmp%d" % os.getpid(), dp + ".tmp%d" % os.getpid()
    t0 = time.time()
    info = build_synth_index(tmp, TX_SEED, genes, 0, device="cuda" if use_gpu else "cpu", verbose=False, perfect_dir=tmpp)
    os.makedirs(os.path.dirname(d), exist_ok=True)
    for name in ("sa.bin", "txpInfo.bin", "rsd.bin"):  # the links of the -p flavour point at the dense directory's final place
        os.remove(os.path.join(tmpp, name))
        os.symlink(os.path.join("..", "idx", name), os.path.join(tmpp, name))
    for a, b in ((tmp, d), (tmpp, dp)):
        try:
            os.rename(a, b)
        except OSError:
            pass
    log(f"built dense + perfect-hash index for {genes} genes in {time.time() - t0:.1f}s ({info['kmers']} k-mers) -> {d}")
    return want + "/"


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
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", os.environ.get("RAPMAP_BENCH_CLOCK_MS", "20"), "-i", str(self.gpu)],
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


def write_fastq_prefix(tx, total: int, d: str):
    """First `total` pairs of the benchmark read stream as FASTQ (for the reference binary), cached."""
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
    return f1, f2


def reference_quasimap(ref_bin, idx, f1, f2, cores, flags):
    """`rapmap_ref quasimap -n`: (mapping seconds from the reference's own ScopedTimer, wall seconds incl. index load)."""
    t0 = time.time()
    p = subprocess.run([ref_bin, "quasimap", "-i", idx, "-1", f1, "-2", f2, "-t", str(cores), "-n"] + flags, capture_output=True, text=True)
    wall = time.time() - t0
    m = re.findall(r"Elapsed time: ([0-9.eE+-]+)s", p.stdout + p.stderr)
    return (float(m[-1]) if m else wall), wall


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
    f1, f2 = write_fastq_prefix(tx, total, d)
    flags = ["-s"] if args.selaln else []
    if args.warmup > 0:  # one untimed pass warms the page cache (the CLI is one-shot: warm-up steps cannot be separated)
        w1, w2 = os.path.join(d, "w1.fastq"), os.path.join(d, "w2.fastq")
        for src, dst in ((f1, w1), (f2, w2)):
            with open(src, "rb") as f, open(dst, "wb") as g:
                g.write(b"".join(f.readline() for _ in range(4 * 2000)))
        reference_quasimap(REF_BIN, idx, w1, w2, cores, flags)
    mapping_s, wall = reference_quasimap(REF_BIN, idx, f1, f2, cores, flags)
    value = total / mapping_s
    line.update({
        "value": value, "ms_per_step": 1e3 * mapping_s / args.steps,
        "config": {"workload": workload_name(args.genes) + (" + -s" if args.selaln else ""),
                   "pairs_per_step": sample, "command": f"rapmap_ref quasimap{' -s' if args.selaln else ''} -n -t {cores}",
                   "timed": "reference ScopedTimer 'Elapsed time' around mapReads: FASTQ parsing (page cache warm) + mapping, index load excluded",
                   "wall_s_incl_index_load": wall,
                   "note": "each step is a bounded sample of the workload (the CPU needs ~12 s for one 10M-pair pass)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": f"{total} pairs (first pairs of the benchmark read stream)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1 << 20, help="pairs per chunk (one rapmap_cuda_map_batch call) per GPU")
    ap.add_argument("--chunks", type=int, default=10, help="chunks per step per GPU (10 x 2^20 = the 10M-read configuration)")
    ap.add_argument("--genes", type=int, default=37000, help="synthetic genes (37000 -> ~203k transcripts)")
    ap.add_argument("--selaln", action="store_true", help="headline leg with quasimap -s (configs[2]) instead of the default flags")
    ap.add_argument("--legs", default="selaln,perfect_hash", help="extra legs in the JSON line (comma separated; 'none' to skip)")
    ap.add_argument("--leg-steps", type=int, default=0, help="steps of the extra legs (0: min(--steps, 10))")
    ap.add_argument("--distinct", type=int, default=4, help="distinct read chunks cycled through the steps")
    ap.add_argument("--ref-pairs-per-step", type=int, default=100000, help="pairs per step of the reference arm (bounded sample)")
    ap.add_argument("--oracle-sample", type=int, default=20000, help="pairs checked against / counted by the CPU oracle at N=1")
    ap.add_argument("--cpu-baseline-pairs", type=int, default=1000000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-in", default="host", choices=["host", "device"], help="diagnosis only: where the end-to-end leg's inputs live")
    ap.add_argument("--e2e-out", default="host", choices=["host", "device"], help="diagnosis only: where the end-to-end leg's outputs go")
    ap.add_argument("--e2e-depth", type=int, default=3, choices=[1, 2, 3], help="chunks the end-to-end leg keeps in flight on its ONE mapper (the mapper pipelines two)")
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

    def max_over_ranks(ms: float) -> float:
        if world > 1:
            tm = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            return float(tm.item())
        return ms

    legs_wanted = [] if args.legs in ("", "none") else [x for x in args.legs.split(",") if x]
    if world > 1:
        legs_wanted = [x for x in legs_wanted if x == "selaln"]  # configs[4]: the -s scaling curve
    if args.selaln:
        legs_wanted = [x for x in legs_wanted if x != "selaln"]

    # ---- index: rank 0 builds (cached) + loads from the reference-format files; other ranks receive the packed image over NCCL
    from rapmap_b200.sharding import replicate_index

    def load_index(perfect: bool):
        if rank == 0:
            ensure_index(args.genes, True, perfect)
        barrier()
        d = index_dir(args.genes, perfect) + "/"
        t0 = time.time()
        ix = rb.Index(d, local) if rank == 0 else None
        ix, keep = replicate_index(ix, rank, local)
        log(f"rank {rank}: {'-p ' if perfect else ''}index ready in {time.time() - t0:.1f}s ({ix.device_bytes / 2**30:.2f} GiB in HBM, {ix.num_transcripts} transcripts)")
        return d, ix, keep

    idx_dir, index, _keep0 = load_index(False)

    # ---- reads: rank-sharded contiguous ranges of the counter-based stream; pinned host copies + device copies
    B, CH = args.batch, max(1, args.chunks)
    tx = SynthTxome(TX_SEED, args.genes)
    nd = max(1, args.distinct)
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
    nm = args.e2e_depth
    d_res = [(torch.empty(cap * 28, dtype=torch.uint8, device="cuda"), torch.empty(B + 1, dtype=torch.int64, device="cuda")) for _ in range(nm)]
    h_out = [(torch.empty(cap * 28, dtype=torch.uint8).pin_memory(), torch.empty(B + 1, dtype=torch.int64).pin_memory()) for _ in range(nm)]
    d_outs = d_res
    dev = torch.device("cuda", local)

    STAGES = ("pack", "sa", "map", "merge", "selaln", "ksw", "h2d", "d2h")

    def run_leg(ix, opts, steps, warmup, sample_clocks=False):
        """Resident leg (device buffers in and out, one mapper, CUDA events on its compute stream) and end-to-end leg (pinned host buffers
        in and out through rapmap_cuda_map_batch_async / _wait: ONE host thread, ONE mapper, two chunks in flight, so one chunk's
        PCIe copies overlap another chunk's kernels) over the same chunk sequence; the two legs must produce the same hits."""
        mappers = [rb.Mapper(ix, opts, max_batch=B, max_read_len=READ_LEN)]
        mapper = mappers[0]
        stream = torch.cuda.ExternalStream(mapper.stream_ptr, device=dev)

        def resident_run(first, count, st=None):
            """Device buffers in and out; chunks go through the same asynchronous pipeline (no host gap between chunks)."""
            def collect():
                r = mapper.wait()
                if st is not None:
                    t = mapper.timing()
                    st["pack"] += t.ms_pack_reads; st["sa"] += t.ms_sa_collect; st["map"] += t.ms_hits_to_mappings; st["merge"] += t.ms_merge
                    st["selaln"] += t.ms_sel_aln; st["ksw"] += t.ms_ksw; st["h2d"] += t.ms_h2d; st["d2h"] += t.ms_d2h
                    st["launch"] += t.launches; st["hits"] += r.num_hits; st["retries"] += t.retries; st["dp_jobs"] += t.dp_jobs; st["dp_jobs_general"] += t.dp_jobs_general

            for c in range(first, first + count):
                if mapper.in_flight == nm:
                    collect()
                a, b = devb[c % nd]
                mapper.map_batch_async(a, b, n=B, fixed_len=READ_LEN, location=rb.LOC_DEVICE, hits_out=d_res[c % nm][0], offsets_out=d_res[c % nm][1],
                                       out_location=rb.LOC_DEVICE, capacity=cap)
            while mapper.in_flight:
                collect()

        resident_run(0, warmup * CH)
        torch.cuda.synchronize()
        barrier()
        sampler = ClockSampler(local) if sample_clocks else None
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st = {k: 0.0 for k in STAGES}
        st.update({"launch": 0, "hits": 0, "retries": 0, "dp_jobs": 0, "dp_jobs_general": 0})
        e0.record(stream)
        resident_run(warmup * CH, steps * CH, st)
        e1.record(stream)
        torch.cuda.synchronize()
        barrier()
        clocks = sampler.stop() if sampler else None
        ms_res = max_over_ranks(e0.elapsed_time(e1))

        # ---- end to end
        e2e_st = {"h2d": 0.0, "d2h": 0.0, "total": 0.0, "sa": 0.0, "pack": 0.0, "map": 0.0, "merge": 0.0, "n": 0}

        def e2e(first, count):
            """ONE host thread, ONE mapper: chunk c+1 is enqueued (rapmap_cuda_map_batch_async) before chunk c is collected
            (rapmap_cuda_mapper_wait), so the mapper's three streams overlap copy-in, kernels and copy-out of consecutive chunks."""
            hits = 0

            def collect():
                nonlocal hits
                hits += mapper.wait().num_hits
                t = mapper.timing()
                e2e_st["h2d"] += t.ms_h2d; e2e_st["d2h"] += t.ms_d2h; e2e_st["total"] += t.ms_total; e2e_st["sa"] += t.ms_sa_collect; e2e_st["n"] += 1
                e2e_st["pack"] += t.ms_pack_reads; e2e_st["map"] += t.ms_hits_to_mappings; e2e_st["merge"] += t.ms_merge

            for c in range(first, first + count):
                if mapper.in_flight == nm:
                    collect()
                a, b = host[c % nd] if args.e2e_in == "host" else devb[c % nd]
                ho, oo = h_out[c % nm] if args.e2e_out == "host" else d_outs[c % nm]
                mapper.map_batch_async(a, b, n=B, fixed_len=READ_LEN, location=rb.LOC_HOST if args.e2e_in == "host" else rb.LOC_DEVICE, hits_out=ho, offsets_out=oo,
                                       out_location=rb.LOC_HOST if args.e2e_out == "host" else rb.LOC_DEVICE, capacity=cap)
            while mapper.in_flight:
                collect()
            return hits

        e2e(0, max(warmup * CH, 4))
        torch.cuda.synchronize()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        for kk in e2e_st:
            e2e_st[kk] = 0
        e2e_hits = e2e(warmup * CH, steps * CH)
        f1.record(stream)  # every chunk has been waited for: the event marks the end of the host-visible work
        torch.cuda.synchronize()
        barrier()
        ms_e2e = max_over_ranks(f0.elapsed_time(f1))
        if e2e_hits != st["hits"]:
            raise SystemExit(f"end-to-end leg produced {e2e_hits} hits, resident leg {st['hits']} over the same chunks; refusing to report a number")
        e2e_chunk = {kk: (e2e_st[kk] / max(1, e2e_st["n"])) for kk in ("h2d", "pack", "sa", "map", "merge", "d2h", "total")}
        return {"ms_res": ms_res, "ms_e2e": ms_e2e, "st": st, "e2e_chunk": e2e_chunk, "clocks": clocks, "mapper": mapper, "mappers": mappers, "steps": steps}

    def oracle_check(o_idx_dir, opts, mapper, ns):
        """Parity + operation counts on the first `ns` pairs of chunk 0 (CPU oracle; N == 1 only)."""
        a, b = host[0][0].numpy()[:ns].copy(), host[0][1].numpy()[:ns].copy()
        om = OracleMapper(o_idx_dir, opts)
        t0 = time.time()
        ref = om.map(a, b, READ_LEN)
        port_s = time.time() - t0
        got = mapper.map_batch(a, b, n=ns, fixed_len=READ_LEN)
        ok = bool(got.num_hits == ref.num_hits and np.array_equal(got.hits, ref.hits) and np.array_equal(got.pair_offsets, ref.pair_offsets)
                  and np.array_equal(got.counters, ref.counters))
        return ok, om.op_counts(), port_s

    def cpu_baseline(flags, ns):
        if args.no_cpu_baseline or not os.path.exists(REF_BIN):
            return None
        cores = os.cpu_count() or 1
        f1, f2 = write_fastq_prefix(tx, ns, os.path.join(CACHE, "cpu_reads"))
        s, _ = reference_quasimap(REF_BIN, idx_dir, f1, f2, cores, flags)
        return {"value": ns / s, "unit": UNIT, "cores": cores, "kind": "reference",
                "sample": f"first {ns} pairs of the benchmark stream, rapmap_ref quasimap{' ' + ' '.join(flags) if flags else ''} -n -t {cores}, reference ScopedTimer (FASTQ parsing + mapping, index load excluded)"}

    def leg_numbers(L, total_pairs_per_step):
        steps = L["steps"]
        return total_pairs_per_step * steps / (L["ms_res"] / 1e3), total_pairs_per_step * steps / (L["ms_e2e"] / 1e3)

    def roofline(L, ops, ns, selaln):
        peak, peak_src = measured_peaks()
        nchunks = L["steps"] * CH
        sa_ms = L["st"]["sa"] / nchunks
        if ops is not None:
            alg = alg_bytes_per_pair(ops, ns)
        else:  # counted on this workload by the instrumented oracle
            alg = (2 * READ_LEN + 16.0 * 121.9 + 4.0 * 61.1 + 792.0) if selaln else (2 * READ_LEN + 16.0 * 96.98 + 4.0 * 17.55 + 648.8)
        achieved = alg * B / (sa_ms / 1e3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "sa_collect_traffic_selaln.json" if selaln else "sa_collect_traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        return {"bound": "hbm", "kernel": "sa_collect_lane_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "alg_bytes_per_pair": alg, "pairs_per_launch": B, "kernel_ms_per_launch": sa_ms,
                "stage_ms_per_chunk": {k: L["st"][k] / nchunks for k in STAGES},
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel at this batch size (profiles/)" if traffic else None}

    pairs_per_step = B * CH * world
    # ---- headline leg
    opts = rb.default_opts(sel_aln=args.selaln)
    H = run_leg(index, opts, args.steps, args.warmup, sample_clocks=True)
    value, e2e_value = leg_numbers(H, pairs_per_step)
    log(f"rank {rank}: headline: resident {H['ms_res'] / args.steps:.2f} ms/step, end to end {H['ms_e2e'] / args.steps:.2f} ms/step (one mapper, one host thread, {nm} chunks in flight)")
    parity, ops, ops_per_pair, ns = None, None, None, 0
    if rank == 0 and world == 1 and args.oracle_sample > 0:
        ns = min(args.oracle_sample, B)
        parity, ops, port_s = oracle_check(idx_dir, opts, H["mapper"], ns)
        ops_per_pair = {k: v / ns for k, v in ops.items()}
        log(f"oracle sample: {ns} pairs in {port_s:.1f}s, parity={parity}, ops/pair={ops_per_pair}")
        if not parity:
            raise SystemExit("PARITY FAILURE against the CPU oracle on the benchmark sample; refusing to report a number")
    headline_hits_chunk0 = None
    if "perfect_hash" in legs_wanted:  # kept for the -p leg: full-chunk results of the dense index
        r0 = H["mapper"].map_batch(host[0][0], host[0][1], n=B, fixed_len=READ_LEN, hits_out=h_out[0][0], offsets_out=h_out[0][1], capacity=cap)
        headline_hits_chunk0 = (r0.num_hits, h_out[0][0][: r0.num_hits * 28].clone(), h_out[0][1].clone())
    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": H["ms_res"] / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {"workload": workload_name(args.genes) + (" + -s" if args.selaln else ""), "transcripts": index.num_transcripts,
                       "pairs_per_step_per_gpu": B * CH, "chunks_per_step": CH, "pairs_per_chunk": B, "distinct_chunks": nd,
                       "l2_policy": f"inputs larger than L2 ({2 * B * READ_LEN / 2**20:.0f} MiB of read bases per chunk, {index.device_bytes / 2**30:.1f} GiB index)",
                       "index": "replicated per GPU (one NCCL broadcast of the packed image)", "sharding": "contiguous read ranges per rank, no data-path collective"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * B * READ_LEN * CH, "d2h_bytes_per_step": int(H["st"]["hits"] / args.steps * 28 + (B + 1) * 8 * CH),
                    "ms_per_step": H["ms_e2e"] / args.steps, "api": "rapmap_cuda_map_batch_async + rapmap_cuda_mapper_wait with pinned HOST buffers (ASCII bases in, rapmap_hit_t records + offsets out)",
                    "host_threads": 1, "mappers": 1, "chunks_in_flight": nm, "hits_equal_resident_leg": True,
                    "per_chunk_ms_on_its_streams": H["e2e_chunk"]},
            "gpu_launches": int(H["st"]["launch"]),
            "clocks": H["clocks"],
            "roofline": roofline(H, ops, ns, args.selaln),
            "cpu_baseline": cpu_baseline(["-s"] if args.selaln else [], args.cpu_baseline_pairs) if world == 1 else None,
            "hits_per_pair": H["st"]["hits"] / (B * CH * args.steps),
            "ops_per_pair": ops_per_pair,
            "parity_checked_vs_oracle": parity,
            "legs": {},
        }
    for mp in H["mappers"]:
        mp.close()
    del H

    # ---- extra legs
    leg_steps = args.leg_steps or min(args.steps, 10)
    for name in legs_wanted:
        if name == "selaln":
            lopts = rb.default_opts(sel_aln=True)
            L = run_leg(index, lopts, leg_steps, args.warmup)
            v, e = leg_numbers(L, pairs_per_step)
            log(f"rank {rank}: leg selaln: resident {L['ms_res'] / leg_steps:.2f} ms/step, end to end {L['ms_e2e'] / leg_steps:.2f} ms/step")
            if rank == 0:
                lp, lops, lns = None, None, 0
                if world == 1 and args.oracle_sample > 0:
                    lns = min(args.oracle_sample, B)
                    lp, lops, _ = oracle_check(idx_dir, lopts, L["mapper"], lns)
                    if not lp:
                        raise SystemExit("PARITY FAILURE (-s) against the CPU oracle on the benchmark sample; refusing to report a number")
                nchunks = leg_steps * CH
                ksw_s = L["st"]["ksw"] / 1e3
                cells_per_job = (lops["kswCells"] / max(1, lops["kswCalls"])) if lops else None
                line["legs"]["selaln"] = {
                    "config": "configs[2]" + (" / configs[4] (scaling)" if world > 1 else "") + ": same index and reads, quasimap -s (ksw2 selective alignment on)",
                    "value": v, "unit": UNIT, "ms_per_step": L["ms_res"] / leg_steps, "steps": leg_steps, "n_gpus": world,
                    "e2e": {"value": e, "unit": UNIT, "ms_per_step": L["ms_e2e"] / leg_steps, "host_threads": 1, "mappers": 1, "chunks_in_flight": nm, "hits_equal_resident_leg": True},
                    "roofline": roofline(L, lops, lns, True),
                    "ksw": {"dp_jobs_per_pair": L["st"]["dp_jobs"] / (B * nchunks), "dp_jobs_per_s": L["st"]["dp_jobs"] / ksw_s if ksw_s > 0 else None,
                            "general_kernel_share": L["st"]["dp_jobs_general"] / max(1, L["st"]["dp_jobs"]),
                            "band_cells_per_job": cells_per_job,
                            "cell_updates_per_s": (L["st"]["dp_jobs"] * cells_per_job / ksw_s) if (cells_per_job and ksw_s > 0) else None,
                            "note": "band cells = sum over anti-diagonals of (en - st + 1) of ksw_extz2_sse41's loop bounds (counted by the oracle on the sample); time = the two ksw kernels, CUDA events"},
                    "hits_per_pair": L["st"]["hits"] / (B * nchunks),
                    "gpu_launches": int(L["st"]["launch"]),
                    "parity": lp, "parity_method": f"records, offsets and counters of the first {lns} pairs identical to the CPU oracle (-s)" if lp is not None else None,
                    "cpu_baseline": cpu_baseline(["-s"], args.cpu_baseline_pairs // 2) if world == 1 else None,
                }
            for mp in L["mappers"]:
                mp.close()
            del L
        elif name == "perfect_hash":
            pdir, pindex, _keep1 = load_index(True)
            lopts = rb.default_opts()
            L = run_leg(pindex, lopts, leg_steps, args.warmup)
            v, e = leg_numbers(L, pairs_per_step)
            log(f"rank {rank}: leg perfect_hash: resident {L['ms_res'] / leg_steps:.2f} ms/step, end to end {L['ms_e2e'] / leg_steps:.2f} ms/step")
            if rank == 0:
                r1 = L["mapper"].map_batch(host[0][0], host[0][1], n=B, fixed_len=READ_LEN, hits_out=h_out[0][0], offsets_out=h_out[0][1], capacity=cap)
                n0, hh, oo = headline_hits_chunk0
                same = bool(r1.num_hits == n0 and torch.equal(h_out[0][0][: n0 * 28], hh) and torch.equal(h_out[0][1], oo))
                if not same:
                    raise SystemExit("PARITY FAILURE: the -p index does not give the dense index's hits; refusing to report a number")
                nchunks = leg_steps * CH
                line["legs"]["perfect_hash"] = {
                    "config": f"configs[3]: the same {index.num_transcripts}-transcript transcriptome indexed with -p (BooPHF minimum perfect hash + FrugalBooMap), same reads, quasimap default flags",
                    "value": v, "unit": UNIT, "ms_per_step": L["ms_res"] / leg_steps, "steps": leg_steps, "n_gpus": world,
                    "e2e": {"value": e, "unit": UNIT, "ms_per_step": L["ms_e2e"] / leg_steps, "host_threads": 1, "mappers": 1, "chunks_in_flight": nm, "hits_equal_resident_leg": True},
                    "sa_lookup_ms_per_launch": L["st"]["sa"] / nchunks,
                    "stage_ms_per_chunk": {k: L["st"][k] / nchunks for k in STAGES},
                    "lookup": "dense table derived at load time from FrugalBooMap::find over every MPHF slot (answers every query exactly as the walk; 4.3 GB of HBM)",
                    "index_image_bytes": pindex.device_bytes, "dense_index_image_bytes": index.device_bytes,
                    "index_files": "hash_info.bph / hash_info.val written by tools/build_index.py (byte-identical to `quasiindex -p` output on the test transcriptomes, tests/test_index_builder.py)",
                    "gpu_launches": int(L["st"]["launch"]),
                    "parity": same, "parity_method": f"all {B} pairs of chunk 0: hit records and offsets byte-identical to the dense index's (which is checked against the CPU oracle)",
                }
            for mp in L["mappers"]:
                mp.close()
            del L, pindex, _keep1
            # the same leg with the BooPHF walked on the device for every lookup (RAPMAP_B200_PHF=walk: no derived table)
            os.environ["RAPMAP_B200_PHF"] = "walk"
            try:
                _pd, windex, _keep2 = load_index(True)
            finally:
                del os.environ["RAPMAP_B200_PHF"]
            wsteps = min(leg_steps, 3)
            L = run_leg(windex, rb.default_opts(), wsteps, args.warmup)
            v, e = leg_numbers(L, pairs_per_step)
            log(f"rank {rank}: leg perfect_hash (walk): resident {L['ms_res'] / wsteps:.2f} ms/step, end to end {L['ms_e2e'] / wsteps:.2f} ms/step")
            if rank == 0:
                line["legs"]["perfect_hash"]["walk"] = {"value": v, "unit": UNIT, "e2e": {"value": e, "unit": UNIT}, "steps": wsteps,
                                                        "sa_lookup_ms_per_launch": L["st"]["sa"] / (wsteps * CH), "index_image_bytes": windex.device_bytes,
                                                        "lookup": "BooPHF level walk + FrugalBooMap verification per lookup on the device (RAPMAP_B200_PHF=walk)"}
            for mp in L["mappers"]:
                mp.close()
            del L, windex, _keep2
        else:
            raise SystemExit(f"unknown leg {name}")

    if rank == 0:
        emit(line)
        log(f"rank 0: {value / 1e6:.1f} M pairs/s resident, {e2e_value / 1e6:.1f} M pairs/s end to end on {world} GPU(s); legs: "
            + ", ".join(f"{k} {v['value'] / 1e6:.1f}/{v['e2e']['value'] / 1e6:.1f}" for k, v in line["legs"].items()))
    barrier()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
