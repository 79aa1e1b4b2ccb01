export RAPMAP_B200_K1=regroup256
NAME=prof_rg256 KERNELS="sa_collect_regroup" SKIP=1 COUNT=1 bash scripts/gpu_prof.sh
