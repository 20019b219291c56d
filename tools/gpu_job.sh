mkdir -p gpurun_out
for mc in 8192 512 1024 2048 4096 16384; do
python bench.py --algo kdt --n 10000000 --dim 128 --metric L2 --rank-dim 16 --maxcheck $mc --in-flight 2 --steps 10 > gpurun_out/r02_bench_kdt_10m128_mc$mc.json 2> gpurun_out/r02_bench_kdt_10m128_mc$mc.err; echo "mc $mc rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_kdt_10m128_mc$mc.json').read().strip().splitlines()[-1]);print('C3 mc $mc', round(d['value']), round(d['e2e']['value']), 'recall', d['recall_at_10'], 'frac', round(d['roofline']['frac'],3), 'pipelined', round(d['pipelined']['value']), 'cpu', round(d['cpu_baseline']['value']), d['parity_vs_reference']['identical_id_lists'], '/', d['parity_vs_reference']['queries_compared'])"
done
