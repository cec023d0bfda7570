#!/bin/bash
# A/B of library variants / env switches on the 1000-column bench:  tools/gpu_ab.sh TAG "ENV=.. ENV=.." ...   (first word of each config = tag)
for cfg in "$@"; do
  tag=${cfg%% *}; envs=${cfg#* }
  for m in ${AB_METHODS:-euler}; do
    env $envs HHD_VERBOSE=1 python bench.py --method $m --steps 3 --warmup 3 --window 1000 --no-cpu-baseline ${AB_ARGS} > gpurun_out/ab_${tag}_$m.json 2> gpurun_out/ab_${tag}_$m.err || tail -5 gpurun_out/ab_${tag}_$m.err
    grep "k_neuron:" gpurun_out/ab_${tag}_$m.err | head -1
  done
done
python tools/benchline.py gpurun_out/ab_*.json
