#!/bin/bash
# the driver's bench invocation (both arms), as the round-end run does it
mkdir -p gpurun_out
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.log; echo "ref exit $?"
SECONDS=0; timeout 1800 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.log; echo "bench exit $?"
echo "bench wall ${SECONDS}s"; grep -E "\[bench\]" gpurun_out/bench_n1.log | tail -20
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/bench_n1.json"))
    print("headline", round(d["value"]/1e6,2), "e2e", round(d["e2e"]["value"]/1e6,2), "frac", round(d["roofline"]["frac"],4), {k:round(v,2) for k,v in d["roofline"]["stage_ms_per_chunk"].items()}, d["cpu_baseline"] and round(d["cpu_baseline"]["value"]/1e6,3), d["parity_checked_vs_oracle"], d["clocks"])
    for k,v in d["legs"].items():
        print(k, round(v["value"]/1e6,2), "e2e", round(v["e2e"]["value"]/1e6,2), v.get("parity"), v.get("ksw"), {a:round(b,2) for a,b in (v.get("stage_ms_per_chunk") or v["roofline"]["stage_ms_per_chunk"]).items()})
except Exception as e:
    print("ERR", e); print(open("gpurun_out/bench_n1.log").read()[-3000:])
PY
