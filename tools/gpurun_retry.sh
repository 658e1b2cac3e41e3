#!/bin/bash
# usage: tools/gpurun_retry.sh [--gpus N] <timeout> <command>: retries while the pod answers "transient" (nothing charged)
GP=""
if [ "$1" == "--gpus" ]; then GP="--gpus $2"; shift 2; fi
T=$1; shift
for i in 1 2 3 4 5 6 7 8 9 10; do
  OUT=$(/usr/local/graft/bin/gpurun $GP --timeout $T -- "$@" 2>&1)
  echo "$OUT" | tail -60
  if echo "$OUT" | grep -q "status=transient"; then echo "[retry $i] pod busy, sleeping"; sleep 150; else break; fi
done
