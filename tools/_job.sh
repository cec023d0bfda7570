mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 60 python bench.py --steps 1 --warmup 3 --window 300 --no-cpu-baseline --no-extras --distinct 25 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k.err || tail -5 gpurun_out/r2k.err
python tools/benchline.py gpurun_out/r2k_bench.json
