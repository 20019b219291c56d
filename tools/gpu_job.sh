mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_refine.py tests/test_capi_symbols.py -q -x --timeout 600 > gpurun_out/r02_gpu_rebuild.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r02_gpu_rebuild.log
