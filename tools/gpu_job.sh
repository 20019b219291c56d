mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_refine.py tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "refine or quantiz" > gpurun_out/r02_gpu_qrefine.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r02_gpu_qrefine.log
