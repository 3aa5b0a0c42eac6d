#!/bin/bash
# usage: tools/_gpuretry.sh <timeout> [--gpus N] -- retries while the pod answers busy (exit code 3)
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $T "$@" -- 'bash tools/_batch.sh' > /tmp/gpurun_last.log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "rc=$rc"; tail -20 /tmp/gpurun_last.log; exit $rc; fi
  sleep 60
done
echo "gave up"; exit 3
