mkdir -p gpurun_out
P="python bench.py --n 2000000 --dim 100 --quantizer opq --raw-type int8 --pq-m 50"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_kernel -c 1 -f -o gpurun_out/prof_r02_pq $P --steps 1 --warmup 1 --no-cpu-baseline --no-clocks > gpurun_out/ncu_r02_pq.log 2>&1; echo "ncu pq rc=$?"
$P --in-flight 2 > gpurun_out/r02_bench_opq_2m100.json 2> gpurun_out/r02_bench_opq_2m100.err; echo "pq 2m rc=$?"
python bench.py --n 100000000 --dim 100 --quantizer opq --raw-type int8 --pq-m 50 --in-flight 2 --steps 10 > gpurun_out/r02_bench_opq_100m100.json 2> gpurun_out/r02_bench_opq_100m100.err; echo "pq 100m rc=$?"
for f in r02_bench_opq_2m100 r02_bench_opq_100m100; do python -c "
import json;d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]);print('$f', round(d['value']), round(d['e2e']['value']), 'recall', d['recall_at_10'], d.get('pipelined',{}).get('value'), d['cpu_baseline']['value'] if d['cpu_baseline'] else None, d['parity_vs_reference'], d['roofline'].get('l2_gather'))"; done
