mkdir -p gpurun_out
timeout 100 python bench.py --steps 2 --warmup 3 --window 500 --no-cpu-baseline --no-extras --distinct 50 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j.err || tail -5 gpurun_out/r2j.err
python tools/benchline.py gpurun_out/r2j_bench.json
