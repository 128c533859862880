#!/bin/bash
# queue-kernel tuning sweep: CTAs per SM x chunk length (run on the GPU box)
for c in 3 4 5; do for k in 10 20 25 50; do
  echo -n "ctas_per_sm=$c chunk=$k : "
  PG_QUEUE_CTAS_PER_SM=$c PG_QUEUE_CHUNK=$k python tools/sweep_rollout.py 65536 2>&1 | tail -1
done; done
