set -x
nproc
timeout 900 python tools/longrun_study.py --engine gpu --seeds 0,1,2,3,4,5,6 --method rk4 --out gpurun_out/r2_longrun_rk4.json > gpurun_out/r2_longrun_rk4.log 2>&1
timeout 900 python tools/longrun_study.py --engine gpu --seeds 0,1,2,3,4 --method gsl_rk2 --out gpurun_out/r2_longrun_gslrk2.json > gpurun_out/r2_longrun_gslrk2.log 2>&1
python bench.py > gpurun_out/r2_bench_default_a.json 2> gpurun_out/r2_bench_default_a.err
tail -3 gpurun_out/r2_longrun_rk4.log gpurun_out/r2_longrun_gslrk2.log
