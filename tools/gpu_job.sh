mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r02_gpu_p.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_gpu_p.log
B="python bench.py --n 1000000 --dim 128 --metric L2 --rank-dim 16 --no-cpu-baseline --no-clocks --steps 10"
$B > gpurun_out/x_v9.json 2> gpurun_out/x_v9.err; python -c "import json;d=json.loads(open('gpurun_out/x_v9.json').read().strip().splitlines()[-1]);print('v9 bkt', round(d['value']),d['roofline']['kernel_ms'])"
$B --nq 40000 --param B200.QueriesPerSM=16 > gpurun_out/x_v9b.json 2> gpurun_out/x_v9b.err; python -c "import json;d=json.loads(open('gpurun_out/x_v9b.json').read().strip().splitlines()[-1]);print('v9 bkt nq40k/16', round(d['value']),d['roofline']['kernel_ms'])"
