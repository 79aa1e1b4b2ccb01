cd /tmp; /root/repo/build/bin/synth reads --genes 8 --seed 777 --pairs 1500 --rseed 4242 --sub 20000 --ins 2000 --del 2000 --n 2000 --out1 y1.fq --out2 y2.fq
for i in 1 2 3 4 5 6 7 8; do /root/repo/oracle/_ref/adapter_sam /root/repo/tests/golden/synth_idx/ /tmp/a$i.sam -1 y1.fq -2 y2.fq --chunk 700 2>/dev/null; md5sum /tmp/a$i.sam; done
for i in 1 2 3 4; do /root/repo/oracle/_ref/adapter_sam /root/repo/tests/golden/synth_idx/ /tmp/b$i.sam -1 y1.fq -2 y2.fq --chunk 10000 2>/dev/null; md5sum /tmp/b$i.sam; done
mkdir -p /root/repo/gpurun_out; cp /tmp/a1.sam /root/repo/gpurun_out/flaky_a1.sam; for i in 2 3 4 5 6 7 8; do cmp -s /tmp/a1.sam /tmp/a$i.sam || { cp /tmp/a$i.sam /root/repo/gpurun_out/flaky_other.sam; break; }; done
