set -x
(CFDP_FLUX_VARIANT=3 timeout 900 python -m pytest tests/test_gpu_flux.py -x -q 2>&1 | tail -3) > gpurun_out/r02b_t4.log 2>&1
KBENCH_FLUX=0,1,2,3 timeout 900 python tools/kbench.py --mpoints 16 --rounds 5 256:lex:0/2.8.0.1 > gpurun_out/r02b_kb4.log 2>&1
