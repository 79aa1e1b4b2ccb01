for cfg in "--e2e-in host --e2e-out host" "--e2e-in device --e2e-out host" "--e2e-in host --e2e-out device" "--e2e-in device --e2e-out device"; do
python bench.py --steps 5 --warmup 2 --no-cpu-baseline --legs none $cfg 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('$cfg', round(d['value']/1e6,1), round(d['e2e']['value']/1e6,1), {k:round(v,2) for k,v in d['e2e']['per_chunk_ms_on_its_streams'].items()})"
done
