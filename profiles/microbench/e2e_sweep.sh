#!/bin/bash
# e2e of the C2 step from pageable host arrays under different host-copy settings (development tool).
for v in "X=1" "MDAGPU_COPY_THREADS=8" "MDAGPU_COPY_THREADS=16" "MDAGPU_RDF_RAMP_DIV=16" "MDAGPU_RDF_RAMP_DIV=4" "MDAGPU_RDF_PIECE_DIV=8" "MDAGPU_RDF_PIECE_DIV=2"; do
  env $v python bench.py --no-secondary --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', 'dev %.2f ms  pinned %.2f ms  pageable %.2f ms' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['pageable']['ms_per_step']))"
done
nproc; python -c "import os; print(len(os.sched_getaffinity(0)))"
