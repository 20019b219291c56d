mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r02_gpu_g.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02_gpu_g.log
B="python bench.py --n 1000000 --dim 128 --metric L2 --rank-dim 16 --no-cpu-baseline --no-clocks --steps 10"
$B > gpurun_out/x_v3.json 2> gpurun_out/x_v3.err; echo "v3 rc=$?"; python -c "import json;d=json.loads(open('gpurun_out/x_v3.json').read().strip().splitlines()[-1]);print(d['value'],d['roofline']['kernel_ms'])"
$B --algo kdt > gpurun_out/x_v3k.json 2> gpurun_out/x_v3k.err; echo "v3k rc=$?"; python -c "import json;d=json.loads(open('gpurun_out/x_v3k.json').read().strip().splitlines()[-1]);print(d['value'],d['roofline']['kernel_ms'])"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_kernel -c 1 -f -o gpurun_out/prof_r02_128_v3 $B --steps 1 --warmup 1 > gpurun_out/ncu_r02_128_v3.log 2>&1; echo "ncu 128 rc=$?"
