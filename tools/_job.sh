mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
HHD_DEBUG_TIMES=gpurun_out/stamps8p timeout 100 $T bench.py --gpus 8 --steps 3 --warmup 3 --window 1000 --no-cpu-baseline --no-extras --distinct 0 > gpurun_out/mg8p_ring.json 2> gpurun_out/mg8p.err || tail -5 gpurun_out/mg8p.err
python tools/benchline.py gpurun_out/mg8p_ring.json
python tools/step_stamps.py gpurun_out/stamps8p.rank0 gpurun_out/stamps8p.rank5
