mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r02_gpu_k.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_gpu_k.log
B="python bench.py --n 1000000 --dim 128 --metric L2 --rank-dim 16 --no-cpu-baseline --no-clocks --steps 10"
$B > gpurun_out/x_v5.json 2> gpurun_out/x_v5.err; python -c "import json;d=json.loads(open('gpurun_out/x_v5.json').read().strip().splitlines()[-1]);print('v5 bkt', round(d['value']),d['roofline']['kernel_ms'])"
$B --nq 40000 --param B200.QueriesPerSM=16 > gpurun_out/x_v5b.json 2> gpurun_out/x_v5b.err; python -c "import json;d=json.loads(open('gpurun_out/x_v5b.json').read().strip().splitlines()[-1]);print('v5 bkt nq40k/16', round(d['value']),d['roofline']['kernel_ms'])"
$B --algo kdt > gpurun_out/x_v5k.json 2> gpurun_out/x_v5k.err; python -c "import json;d=json.loads(open('gpurun_out/x_v5k.json').read().strip().splitlines()[-1]);print('v5 kdt', round(d['value']),d['roofline']['kernel_ms'])"
P="python bench.py --n 2000000 --dim 100 --quantizer opq --raw-type int8 --pq-m 50 --no-cpu-baseline --no-clocks"
$P --steps 10 > gpurun_out/x_pq0.json 2> gpurun_out/x_pq0.err; python -c "import json;d=json.loads(open('gpurun_out/x_pq0.json').read().strip().splitlines()[-1]);print('pq 2m', round(d['value']),d['roofline']['kernel_ms'])"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_kernel -c 1 -f -o gpurun_out/prof_r02_pq_before $P --steps 1 --warmup 1 > gpurun_out/ncu_r02_pq.log 2>&1; echo "ncu pq rc=$?"
