#!/bin/bash
# usage: tools/gpu_sweep.sh  (runs on the GPU box; writes gpurun_out/sweep_*.json)
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in 5 4 3 2; do
  HHD_NEURON_CTAS_PER_SM=$c python bench.py --columns 1000 --method euler --steps 3 --warmup 3 --window 1000 --no-cpu-baseline > gpurun_out/sweep_euler_c$c.json 2> gpurun_out/sweep_err.log || tail -5 gpurun_out/sweep_err.log
done
for m in rk2 rk4; do
  python bench.py --columns 1000 --method $m --steps 3 --warmup 3 --window 1000 --no-cpu-baseline > gpurun_out/sweep_$m.json 2> gpurun_out/sweep_err.log || tail -5 gpurun_out/sweep_err.log
done
python tools/benchline.py gpurun_out/sweep_*.json
