mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2_auto.json 2> gpurun_out/r02_bench_n2_auto.err; echo "n2 rc=$?"
tail -c 3000 gpurun_out/r02_bench_n2_auto.json; tail -15 gpurun_out/r02_bench_n2_auto.err
timeout 600 python -m pytest tests/test_gpu_group.py -q --timeout 600 2>&1 | tail -3
