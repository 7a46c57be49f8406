#!/bin/bash
# Launch N ranks of a python script on one node WITHOUT torchrun, each under its own
# hard timeout (a hung rank cannot outlive it), one log per rank.
#   tools/run_ranks.sh N TIMEOUT_S LOGPREFIX script.py args...
N=$1; T=$2; LOG=$3; shift 3
export MASTER_ADDR=127.0.0.1 MASTER_PORT=${MASTER_PORT:-29533} WORLD_SIZE=$N
pids=()
for r in $(seq 0 $((N-1))); do
  RANK=$r LOCAL_RANK=$r timeout -s KILL $T python "$@" > ${LOG}.rank$r.log 2>&1 &
  pids+=($!)
done
rc=0
for p in "${pids[@]}"; do wait $p || rc=$?; done
exit $rc
