mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_spann_head.py tests/test_gpu_dropin.py -q --timeout 900 > gpurun_out/r02_gpu_c.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_gpu_c.log
python bench.py > gpurun_out/r02_bench_c2.json 2> gpurun_out/r02_bench_c2.err; echo "bench rc=$?"; tail -c 1800 gpurun_out/r02_bench_c2.json; tail -5 gpurun_out/r02_bench_c2.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_kernel -c 1 -f -o gpurun_out/prof_r02_c2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-clocks > gpurun_out/ncu_r02_c2.log 2>&1; echo "ncu c2 rc=$?"
python bench.py --n 1000000 --dim 128 --metric L2 --rank-dim 16 > gpurun_out/r02_bench_bkt_1m128_before.json 2> gpurun_out/r02_bench_bkt_1m128_before.err; echo "bench128 rc=$?"; tail -c 1500 gpurun_out/r02_bench_bkt_1m128_before.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_kernel -c 1 -f -o gpurun_out/prof_r02_128_before python bench.py --n 1000000 --dim 128 --metric L2 --rank-dim 16 --steps 1 --warmup 1 --no-cpu-baseline --no-clocks > gpurun_out/ncu_r02_128.log 2>&1; echo "ncu 128 rc=$?"
ls -la gpurun_out/*.ncu-rep
