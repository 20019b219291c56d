mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r02_gpu_e.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02_gpu_e.log
python bench.py --n 1000000 --dim 128 --metric L2 --rank-dim 16 --no-cpu-baseline > gpurun_out/r02_bench_bkt_1m128_v1.json 2> gpurun_out/r02_bench_bkt_1m128_v1.err; echo "bench128 rc=$?"; tail -c 900 gpurun_out/r02_bench_bkt_1m128_v1.json
python bench.py --n 1000000 --dim 128 --metric L2 --rank-dim 16 --algo kdt --no-cpu-baseline > gpurun_out/r02_bench_kdt_1m128_v1.json 2> gpurun_out/r02_bench_kdt_1m128_v1.err; echo "benchkdt rc=$?"; tail -c 900 gpurun_out/r02_bench_kdt_1m128_v1.json
