mkdir -p gpurun_out
T=${TAG:-b3}
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r02_gpu_$T.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_gpu_$T.log
show() { python -c "import json,sys;d=json.loads(open('gpurun_out/$1.json').read().strip().splitlines()[-1]);print('$1', round(d['value']),d['roofline']['kernel_ms'], round(d['roofline']['frac'],3), round(d.get('pipelined',{}).get('value',0)), (d.get('parity_vs_reference') or {}).get('identical_distance_bits'))"; }
B="python bench.py --n 1000000 --dim 128 --metric L2 --rank-dim 16 --no-clocks --steps 10 --in-flight 2"
$B > gpurun_out/x_${T}_bkt.json 2> gpurun_out/x_${T}_bkt.err; show x_${T}_bkt
$B --no-cpu-baseline --param B200.QueriesPerSM=17 --param B200.Stages=2 > gpurun_out/x_${T}_bkt_s17st2.json 2> gpurun_out/x_${T}_bkt_s17st2.err; show x_${T}_bkt_s17st2
$B --no-cpu-baseline --param B200.QueriesPerSM=19 > gpurun_out/x_${T}_bkt_s19.json 2> gpurun_out/x_${T}_bkt_s19.err; show x_${T}_bkt_s19
$B --no-cpu-baseline --nq 40000 > gpurun_out/x_${T}_bkt_q40.json 2> gpurun_out/x_${T}_bkt_q40.err; show x_${T}_bkt_q40
$B --algo kdt > gpurun_out/x_${T}_kdt.json 2> gpurun_out/x_${T}_kdt.err; show x_${T}_kdt
python bench.py --n 100000 --dim 128 --metric L2 --rank-dim 16 --no-clocks --steps 10 --in-flight 2 > gpurun_out/x_${T}_bkt100k.json 2> gpurun_out/x_${T}_bkt100k.err; show x_${T}_bkt100k
