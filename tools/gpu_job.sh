# The round's standard GPU job (run through gpurun): the -m gpu suite, smoke(), the headline bench line and its
# reference arm.  The other lines under profiles/ were produced with the commands in DESIGN.md section 5 / this file's
# git history (tools/gpu_job_final*.sh).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/gpu_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
python bench.py --in-flight 2 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"; tail -c 700 gpurun_out/bench_c2.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_c2_reference_arm.json 2> gpurun_out/bench_c2_reference_arm.err; echo "ref arm rc=$?"
