mkdir -p gpurun_out
python bench.py --n 100000000 --dim 100 --quantizer opq --raw-type int8 --pq-m 50 --in-flight 2 --steps 10 > gpurun_out/r02_bench_opq_100m100.json 2> gpurun_out/r02_bench_opq_100m100.err; echo "pq 100m rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_opq_100m100.json').read().strip().splitlines()[-1]);print(round(d['value']), round(d['e2e']['value']), 'recall', d['recall_at_10'], round(d['pipelined']['value']), d['cpu_baseline']['value'], d['parity_vs_reference'], d['roofline']['kernel_ms'])"
tail -3 gpurun_out/r02_bench_opq_100m100.err
