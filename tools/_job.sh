set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -m gpu -x -q ) > gpurun_out/r2b_parity.log 2>&1
tail -4 gpurun_out/r2b_parity.log
B="--steps 3 --warmup 3 --window 1000 --no-cpu-baseline --no-extras"
python bench.py $B > gpurun_out/ab_lsahead2.json 2> gpurun_out/ab.err || tail -5 gpurun_out/ab.err
HHD_LIB=$PWD/rescience_hass_hertag_durstewitz_2016_b200/csrc/libhhd_mb5.so python bench.py $B > gpurun_out/ab_lsahead2_mb5.json 2> gpurun_out/ab.err || tail -5 gpurun_out/ab.err
python bench.py $B --workload columns > gpurun_out/ab_lsahead2_tiled.json 2> gpurun_out/ab.err || tail -5 gpurun_out/ab.err
python tools/benchline.py gpurun_out/ab_*.json
P="--steps 1 --warmup 1 --window 1000 --no-cpu-baseline --no-extras"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^k_(neuron|synapse)' -s 3000 -c 400 --csv --log-file gpurun_out/r2b_launches.csv python bench.py $P > gpurun_out/r2b_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:^k_(neuron|synapse)' -s 3000 -c 4 -o gpurun_out/r2b_kernels python bench.py $P > gpurun_out/r2b_ncu_full.log 2>&1
ls -la gpurun_out
