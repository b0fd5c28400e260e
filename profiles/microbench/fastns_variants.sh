#!/bin/bash
# Builds csrc/nsgrid.cu with different CTA shapes / queue depths on the box and times the C3 search
# (bench.py --configs C3: kernel time from the library's events).  Development tool.
cd mdanalysis_b200/csrc
for v in "-DRP_THREADS_=192 -DRP_MINB_=4" "-DRP_THREADS_=128 -DRP_MINB_=6" "-DRP_THREADS_=192 -DRP_MINB_=4 -DRP_LQ_=16" "-DRP_THREADS_=192 -DRP_MINB_=4 -DRP_LQ_=32" "-DRP_THREADS_=160 -DRP_MINB_=5" $EXTRA_VARIANTS; do
  rm -f nsgrid.o
  make EXTRA="$v" nsgrid.o libmdagpu.so > /dev/null 2>&1 || { echo "build failed: $v"; continue; }
  echo "== $v  $(grep -A3 'row_pairs_kernelILi3ELb1ENS_8PairSink' nsgrid.ptxas.log | grep Used | sed 's/ptxas info    : //')"
  (cd ../..; python bench.py --configs C3 --no-cpu --steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['secondary']['C3_fastns_1M']
print('   triclinic %.3f ms   ortho %.3f ms' % (c['self_search_triclinic']['search_ms'], c['self_search_ortho']['search_ms']))")
done
rm -f nsgrid.o; make nsgrid.o libmdagpu.so > /dev/null 2>&1
