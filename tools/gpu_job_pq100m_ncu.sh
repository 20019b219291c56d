mkdir -p gpurun_out
P="python bench.py --n 100000000 --dim 100 --quantizer opq --raw-type int8 --pq-m 50"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:search_kernel -c 1 -f -o gpurun_out/prof_r02_pq100m $P --steps 1 --warmup 1 --no-cpu-baseline --no-clocks > gpurun_out/ncu_r02_pq100m.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_r02_pq100m.log
