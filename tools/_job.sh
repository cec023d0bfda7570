mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k 'regex:^k_neuron' -s 1500 -c 2 -o gpurun_out/r2i_k_neuron_rk4 python bench.py --steps 1 --warmup 1 --window 1000 --no-cpu-baseline --no-extras --distinct 0 --method rk4 > gpurun_out/r2i_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -2
