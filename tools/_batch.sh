set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
(CFDP_DIRECT=0 $TR --master-port 29551 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e --no-flux --no-cpu --no-parity) > gpurun_out/r04_bench64_2gpu_putnotify.log 2>&1
($TR --master-port 29552 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e --no-flux --no-cpu --no-parity --variant mpi_async) > gpurun_out/r04_bench64_2gpu_nccl.log 2>&1
