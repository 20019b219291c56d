mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r02_gpu_q.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_gpu_q.log
B="python bench.py --n 1000000 --dim 128 --metric L2 --rank-dim 16 --no-cpu-baseline --no-clocks --steps 10 --in-flight 2"
$B > gpurun_out/x_v10.json 2> gpurun_out/x_v10.err; python -c "import json;d=json.loads(open('gpurun_out/x_v10.json').read().strip().splitlines()[-1]);print('v10 bkt', round(d['value']),d['roofline']['kernel_ms'], d['pipelined'])"
$B --algo kdt > gpurun_out/x_v10k.json 2> gpurun_out/x_v10k.err; python -c "import json;d=json.loads(open('gpurun_out/x_v10k.json').read().strip().splitlines()[-1]);print('v10 kdt', round(d['value']),d['roofline']['kernel_ms'], d['pipelined'])"
python bench.py --no-cpu-baseline --no-clocks --steps 10 --in-flight 2 > gpurun_out/x_c2_v10.json 2> gpurun_out/x_c2_v10.err; python -c "import json;d=json.loads(open('gpurun_out/x_c2_v10.json').read().strip().splitlines()[-1]);print('C2', round(d['value']),d['roofline']['kernel_ms'], d['roofline']['frac'], d['pipelined'])"
P="python bench.py --n 2000000 --dim 100 --quantizer opq --raw-type int8 --pq-m 50 --no-cpu-baseline --no-clocks --steps 10 --in-flight 2"
$P > gpurun_out/x_pq_v10.json 2> gpurun_out/x_pq_v10.err; python -c "import json;d=json.loads(open('gpurun_out/x_pq_v10.json').read().strip().splitlines()[-1]);print('pq 2m', round(d['value']),d['roofline']['kernel_ms'], d['pipelined'], d['roofline'].get('l2_gather'))"
