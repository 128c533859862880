#!/bin/bash
# queue-kernel tuning sweep (run on the GPU box): each line of CONFIGS is a set of PG_QUEUE_* assignments
CONFIGS=${CONFIGS:-"PG_QUEUE_CHUNK=20;PG_QUEUE_CMAX=100 PG_QUEUE_CMIN=8;PG_QUEUE_CMAX=50 PG_QUEUE_CMIN=8;PG_QUEUE_CMAX=167 PG_QUEUE_CMIN=8;PG_QUEUE_CMAX=100 PG_QUEUE_CMIN=4;PG_QUEUE_CMAX=100 PG_QUEUE_CMIN=16;PG_QUEUE_CMAX=100 PG_QUEUE_CMIN=8 PG_QUEUE_CTAS_PER_SM=4"}
IFS=';' read -ra CFG <<< "$CONFIGS"
for c in "${CFG[@]}"; do
  echo -n "$c : "
  env $c python tools/sweep_rollout.py ${BATCH:-65536} 2>&1 | tail -1
done
