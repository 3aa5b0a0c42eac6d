set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
$TR --master-port 29541 tools/f6like_configs.py gpu 24 > gpurun_out/r04_config3_f6like24_4gpu.log 2>&1
