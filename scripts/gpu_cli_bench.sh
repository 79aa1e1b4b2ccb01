#!/bin/bash
# FASTQ -> SAM through the CLI front end next to the reference's own CLI on the same files (rows f1/f2 of SURVEY.md section 8)
mkdir -p gpurun_out
PAIRS=${PAIRS:-4000000}
python bench.py --steps 1 --warmup 0 --no-cpu-baseline --oracle-sample 0 --legs none --chunks 1 > /dev/null 2>&1   # builds + caches the index
IDX=/tmp/rapmap_b200_cache/bench_g37000_s12345/idx/
D=/tmp/cli_bench; mkdir -p $D
[ -f $D/r2.fastq ] || build/bin/synth reads --genes 37000 --seed 12345 --pairs $PAIRS --rseed 54321 --out1 $D/r1.fastq --out2 $D/r2.fastq
ls -la $D | tail -2
NP=$(nproc)
{
echo "pairs $PAIRS, host threads $NP, index $IDX"
for mode in "-n" "-o /dev/null"; do
  for fl in "" "-s"; do
    build/bin/rapmap_b200 quasimap -i $IDX -1 $D/r1.fastq -2 $D/r2.fastq -t $NP $mode $fl 2> $D/ours.log; ours=$(grep -o "Elapsed time: [0-9.e+-]*" $D/ours.log | tail -1)
    oracle/_ref/rapmap_ref quasimap -i $IDX -1 $D/r1.fastq -2 $D/r2.fastq -t $NP $mode $fl > $D/ref.log 2>&1; ref=$(grep -o "Elapsed time: [0-9.e+-]*" $D/ref.log | tail -1)
    echo "mode [$mode] flags [$fl]  rapmap_b200: $ours  |  rapmap_ref -t $NP: $ref"
  done
done
echo "second pass, engine only (files and index warm), default wait (poll) | RAPMAP_B200_WAIT=block"
for mode in "-n" "-o /dev/null"; do
  for fl in "" "-s"; do
    build/bin/rapmap_b200 quasimap -i $IDX -1 $D/r1.fastq -2 $D/r2.fastq -t $NP $mode $fl 2> $D/ours.log; ours=$(grep -o "Elapsed time: [0-9.e+-]*" $D/ours.log | tail -1)
    RAPMAP_B200_WAIT=block build/bin/rapmap_b200 quasimap -i $IDX -1 $D/r1.fastq -2 $D/r2.fastq -t $NP $mode $fl 2> $D/ours.log; blk=$(grep -o "Elapsed time: [0-9.e+-]*" $D/ours.log | tail -1)
    echo "mode [$mode] flags [$fl]  rapmap_b200: $ours  |  blocking wait: $blk"
  done
done
} | tee gpurun_out/cli_bench.txt
