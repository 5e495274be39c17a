bash scripts/gpu_exp.sh 0 128 0 128
bash scripts/gpu_trace.sh 0
