bash scripts/gpu_ab.sh
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sa_collect_lane|pack_reads|hits_to_mappings" -s 3 -c 3 -o gpurun_out/prof_r01d -f \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --oracle-sample 0 > /dev/null 2> gpurun_out/ncu_full.log
echo "ncu full exit $?"
