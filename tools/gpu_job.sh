mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r02_gpu_m.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_gpu_m.log
python bench.py --no-cpu-baseline --no-clocks --steps 10 > gpurun_out/x_c2_v6.json 2> gpurun_out/x_c2_v6.err; python -c "import json;d=json.loads(open('gpurun_out/x_c2_v6.json').read().strip().splitlines()[-1]);print('C2', round(d['value']),d['roofline']['kernel_ms'], d['roofline']['frac'])"
python bench.py --n 1000000 --dim 128 --metric L2 --rank-dim 16 --raw-type int8 --no-cpu-baseline --no-clocks --steps 10 > gpurun_out/x_i8_v6.json 2> gpurun_out/x_i8_v6.err; python -c "import json;d=json.loads(open('gpurun_out/x_i8_v6.json').read().strip().splitlines()[-1]);print('int8 1m128', round(d['value']),d['roofline']['kernel_ms'], d['roofline']['frac'])"
P="python bench.py --n 2000000 --dim 100 --quantizer opq --raw-type int8 --pq-m 50 --no-cpu-baseline --no-clocks --steps 10"
$P > gpurun_out/x_pq_v6.json 2> gpurun_out/x_pq_v6.err; python -c "import json;d=json.loads(open('gpurun_out/x_pq_v6.json').read().strip().splitlines()[-1]);print('pq 2m', round(d['value']),d['roofline']['kernel_ms'])"
B="python bench.py --n 1000000 --dim 128 --metric L2 --rank-dim 16 --no-cpu-baseline --no-clocks --steps 10"
$B > gpurun_out/x_v6.json 2> gpurun_out/x_v6.err; python -c "import json;d=json.loads(open('gpurun_out/x_v6.json').read().strip().splitlines()[-1]);print('v6 bkt', round(d['value']),d['roofline']['kernel_ms'])"
$B --algo kdt > gpurun_out/x_v6k.json 2> gpurun_out/x_v6k.err; python -c "import json;d=json.loads(open('gpurun_out/x_v6k.json').read().strip().splitlines()[-1]);print('v6 kdt', round(d['value']),d['roofline']['kernel_ms'])"
