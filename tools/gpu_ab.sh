#!/bin/bash
# A/B of env-var switches on the 1000-column bench: tools/gpu_ab.sh "HHD_PDL=0" "HHD_PDL=1"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $cfg python bench.py --columns 1000 --method euler --steps 3 --warmup 3 --window 1000 --no-cpu-baseline > gpurun_out/ab_$i.json 2> gpurun_out/ab_err.log || tail -5 gpurun_out/ab_err.log
  echo "$cfg"; python tools/benchline.py gpurun_out/ab_$i.json
done
