mkdir -p gpurun_out
python bench.py --builder reference --n 100000 --dim 768 --metric Cosine --in-flight 2 > gpurun_out/r02_bench_refbuilt_100k768.json 2> gpurun_out/r02_bench_refbuilt_100k768.err; echo "refbuilt rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_refbuilt_100k768.json').read().strip().splitlines()[-1]);print('refbuilt', round(d['value']), round(d['e2e']['value']), 'recall', d['recall_at_10'], 'frac', round(d['roofline']['frac'],3), d['cpu_baseline']['value'], d['parity_vs_reference'], d['config']['index_builder'])"
tail -3 gpurun_out/r02_bench_refbuilt_100k768.err
