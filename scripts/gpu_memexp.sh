#!/bin/bash
# memory-system experiment: DRAM / L2 sector counts of the SA-lookup kernel under different load flavours
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 0 --no-cpu-baseline --oracle-sample 0 > /dev/null 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_op_read.sum,lts__t_sector_hit_rate.pct,l1tex__m_xbar2l1tex_read_sectors.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,lts__t_sectors_srcunit_tex_aperture_device_op_read_lookup_miss.sum,smsp__inst_executed.sum
run() { # name env lib
  env $2 RAPMAP_B200_LIB=$3 timeout 300 ncu --metrics $M --clock-control none -k regex:"sa_collect_lane" -s 1 -c 1 --csv --log-file gpurun_out/mem_$1.csv \
     python bench.py --steps 1 --warmup 1 --no-cpu-baseline --oracle-sample 0 --e2e-mappers 1 > /dev/null 2> gpurun_out/mem_$1.log
  echo "== $1"; grep -v "^==" gpurun_out/mem_$1.csv | python -c "
import csv,sys
for r in csv.reader(sys.stdin):
    if len(r)>3 and r[0]!='ID': print('  ', r[-3], r[-1], r[-2])
"
}
run default A=1 ""
for lib in rapmap_b200/_build/ab/lib_*.so; do   # e.g. build_variant.sh nofilter -DRAPMAP_NO_FILTER; build_variant.sh ldcg -DRAPMAP_LDCG
  [ -f "$lib" ] || continue
  n=$(basename $lib .so); run ${n#lib_} A=1 $PWD/$lib
done
