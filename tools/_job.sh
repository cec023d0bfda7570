set -x
mkdir -p gpurun_out
nproc; free -g | head -2; df -h /tmp | tail -1
( time python bench.py --steps 3 --warmup 3 --window 1000 --no-cpu-baseline --no-extras ) > gpurun_out/r2e_distinct.json 2> gpurun_out/r2e_distinct.err || tail -20 gpurun_out/r2e_distinct.err
tail -4 gpurun_out/r2e_distinct.err
python tools/benchline.py gpurun_out/r2e_distinct.json
( time python bench.py --steps 3 --warmup 3 --window 1000 --no-cpu-baseline --no-extras ) > gpurun_out/r2e_distinct_cached.json 2> gpurun_out/r2e_distinct2.err || tail -20 gpurun_out/r2e_distinct2.err
tail -4 gpurun_out/r2e_distinct2.err
python tools/benchline.py gpurun_out/r2e_distinct_cached.json
( time python bench.py --config3-only ) > gpurun_out/r2e_config3.json 2> gpurun_out/r2e_config3.err || tail -20 gpurun_out/r2e_config3.err
cat gpurun_out/r2e_config3.json | cut -c1-1800
du -sh /tmp/hhd_columns | tail -1
