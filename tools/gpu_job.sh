mkdir -p gpurun_out
timeout 900 python tools/rebuild_bench.py --num-vectors 1000000 --dim 128 > gpurun_out/r02_rebuild_1m128.json 2> gpurun_out/r02_rebuild_1m128.err; echo "rc=$?"; tail -1 gpurun_out/r02_rebuild_1m128.json; tail -3 gpurun_out/r02_rebuild_1m128.err
