mkdir -p gpurun_out
T=${TAG:-b4}
show() { python -c "import json,sys;d=json.loads(open('gpurun_out/$1.json').read().strip().splitlines()[-1]);print('$1', round(d['value']),d['roofline']['kernel_ms'], round(d['roofline']['frac'],3), round(d.get('pipelined',{}).get('value',0)))"; }
B="python bench.py --n 1000000 --dim 128 --metric L2 --rank-dim 16 --no-cpu-baseline --no-clocks --steps 10 --in-flight 2 --algo kdt"
for S in 10 12 14 16 17; do
$B --param B200.QueriesPerSM=$S > gpurun_out/x_${T}_kdt_s$S.json 2> gpurun_out/x_${T}_kdt_s$S.err; show x_${T}_kdt_s$S
done
$B --param B200.QueriesPerSM=14 --param B200.NGCacheEntries=600 --param B200.SPTCacheEntries=600 > gpurun_out/x_${T}_kdt_s14h.json 2> gpurun_out/x_${T}_kdt_s14h.err; show x_${T}_kdt_s14h
