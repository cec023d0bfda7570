set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t3.log 2>&1
tail -5 gpurun_out/r2_t3.log
timeout 600 python -m pytest tests/test_gpu_longrun.py -m gpu -x -q -s > gpurun_out/r2_t3_longrun.log 2>&1
tail -30 gpurun_out/r2_t3_longrun.log
