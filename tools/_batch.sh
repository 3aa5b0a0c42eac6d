set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
(time $TR --master-port 29521 bench.py --gpus 8 --mpoints 256 --steps 20 --warmup 5) > gpurun_out/r03_bench256_8gpu.log 2>&1
$TR --master-port 29522 tools/f6like_configs.py gpu 24 > gpurun_out/r03_config3_f6like24_8gpu.log 2>&1
(timeout 600 python -m pytest tests/test_multigpu.py -x -q -k "8-24 or 4-12" 2>&1 | tail -3) > gpurun_out/r03_mg8.log 2>&1
