set -x
(timeout 600 python -m pytest tests/test_gpu_scale.py tests/test_gpu_parity.py -x -q -k "optional_kernel or var_refresh" 2>&1 | tail -5) > gpurun_out/r04_t1.log 2>&1
