mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_gpu_final.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke.log
python bench.py --in-flight 2 > gpurun_out/r02_bench_c2.json 2> gpurun_out/r02_bench_c2.err; echo "bench c2 rc=$?"
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_c2_reference_arm.json 2> gpurun_out/r02_bench_c2_reference_arm.err; echo "ref arm rc=$?"; tail -c 600 gpurun_out/r02_bench_c2_reference_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-clocks > gpurun_out/r02_launches_c2.log 2>&1; echo "launch list rc=$?"
python bench.py --n 1000000 --dim 128 --metric L2 --rank-dim 16 --in-flight 2 > gpurun_out/r02_bench_bkt_1m128.json 2> gpurun_out/r02_bench_bkt_1m128.err; echo "bench 128 rc=$?"
python bench.py --n 1000000 --dim 128 --metric L2 --rank-dim 16 --algo kdt --in-flight 2 > gpurun_out/r02_bench_kdt_1m128.json 2> gpurun_out/r02_bench_kdt_1m128.err; echo "bench kdt 1m rc=$?"
python bench.py --n 100000 --dim 128 --metric L2 --rank-dim 16 --in-flight 2 > gpurun_out/r02_bench_bkt_100k128.json 2> gpurun_out/r02_bench_bkt_100k128.err; echo "bench C1 shape rc=$?"
for f in r02_bench_c2 r02_bench_bkt_1m128 r02_bench_kdt_1m128 r02_bench_bkt_100k128; do python -c "
import json;d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]);print('$f', round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],3), d.get('pipelined',{}).get('value'), d['cpu_baseline']['value'] if d['cpu_baseline'] else None, d['parity_vs_reference'])"; done
