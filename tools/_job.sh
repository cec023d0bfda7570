set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -m gpu -x -q > gpurun_out/r2_t5.log 2>&1
tail -25 gpurun_out/r2_t5.log
AB_METHODS="euler" bash tools/gpu_ab.sh "ring A=1"
AB_METHODS="euler" AB_ARGS="--workload columns" bash tools/gpu_ab.sh "ringtiled A=1"
AB_METHODS="euler" AB_ARGS="--idc-scale 3.0" bash tools/gpu_ab.sh "ringhigh A=1"
