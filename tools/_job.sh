mkdir -p gpurun_out
( time timeout 150 python -m pytest tests/test_gpu_revision.py tests/test_gpu_scaled.py -m gpu -x -q ) > gpurun_out/r2h_newtests.log 2>&1
tail -15 gpurun_out/r2h_newtests.log
timeout 120 python bench.py --steps 3 --warmup 3 --window 1000 --no-cpu-baseline --no-extras --distinct 200 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h.err || tail -5 gpurun_out/r2h.err
python tools/benchline.py gpurun_out/r2h_bench.json
