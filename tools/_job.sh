set -x
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2f_gputests.log 2>&1
tail -6 gpurun_out/r2f_gputests.log
( time timeout 900 python bench.py ) > gpurun_out/r2f_bench_default.json 2> gpurun_out/r2f_bench_default.err || tail -20 gpurun_out/r2f_bench_default.err
tail -3 gpurun_out/r2f_bench_default.err
python tools/benchline.py gpurun_out/r2f_bench_default.json
B="--steps 3 --warmup 3 --window 1000 --no-cpu-baseline --no-extras"
HHD_PDL=1 python bench.py $B > gpurun_out/r2f_pdl.json 2> gpurun_out/r2f.err || tail -5 gpurun_out/r2f.err
python bench.py $B --method rk4 > gpurun_out/r2f_rk4.json 2> gpurun_out/r2f.err || tail -5 gpurun_out/r2f.err
python tools/benchline.py gpurun_out/r2f_pdl.json gpurun_out/r2f_rk4.json
P="--steps 1 --warmup 1 --window 1000 --no-cpu-baseline --no-extras"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^k_(neuron|synapse)' -s 3000 -c 400 --csv --log-file gpurun_out/r2f_launches.csv python bench.py $P > gpurun_out/r2f_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:^k_(neuron|synapse)' -s 3000 -c 4 -o gpurun_out/r2f_kernels python bench.py $P > gpurun_out/r2f_ncu_full.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
