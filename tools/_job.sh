set -x
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_analysis.py tests/test_fI_closed_form.py -m gpu -x -q ) > gpurun_out/r2g_gputests.log 2>&1
tail -4 gpurun_out/r2g_gputests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time timeout 600 python bench.py --no-config3 ) > gpurun_out/r2g_bench_default.json 2> gpurun_out/r2g_bench_default.err || tail -20 gpurun_out/r2g_bench_default.err
python tools/benchline.py gpurun_out/r2g_bench_default.json
