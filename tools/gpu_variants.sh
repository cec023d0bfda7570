#!/bin/bash
# runs bench on each prebuilt kernel variant tools/libhhd_*.so (GPU box)
for v in tools/libhhd_*.so; do
  n=$(basename $v .so)
  HHD_LIB=$PWD/$v python bench.py --columns 1000 --method euler --steps 3 --warmup 3 --window 1000 --no-cpu-baseline > gpurun_out/var_$n.json 2> gpurun_out/var_err.log || tail -5 gpurun_out/var_err.log
done
python bench.py --columns 1000 --method euler --steps 3 --warmup 3 --window 1000 --no-cpu-baseline > gpurun_out/var_default.json 2> gpurun_out/var_err.log || tail -5 gpurun_out/var_err.log
python tools/benchline.py gpurun_out/var_*.json
