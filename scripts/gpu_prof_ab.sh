NAME=prof_r01f bash scripts/gpu_prof.sh
STEPS=5 bash scripts/gpu_ab.sh
