#!/usr/bin/env bash
# usage: tools/gpu_retry.sh <stdout-file> <gpurun args...>   -- retries while the pod answers "transient" (nothing charged)
out=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$out" 2>&1
  if ! grep -q "status=transient" "$out"; then exit 0; fi
  sleep 90
done
exit 3
