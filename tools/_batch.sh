set -x
(time python bench.py) > gpurun_out/r04_bench64.log 2>&1
T="python tools/ncu_target.py 64 lex 2.8.0 6"
$T > gpurun_out/r04_plain_target.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gg_tile_pipe -s 4 -c 2 --csv --log-file gpurun_out/r04_traffic64.csv $T > gpurun_out/r04_ncu_traffic.log 2>&1
