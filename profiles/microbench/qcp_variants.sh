#!/bin/bash
# Builds csrc/qcp.cu with different group widths / chunk sizes / store paths on the box and times
# the streamed QCP kernel at the C4 shape (nvcc is part of the image).  Development tool.
cd mdanalysis_b200/csrc
for v in "-DQS_GT_=32 -DQS_CH_=256" "-DQS_GT_=64 -DQS_CH_=512" "-DQS_GT_=128 -DQS_CH_=1024" "-DQS_GT_=32 -DQS_CH_=256 -DQCP_DF_TRANSLATE=0" "-DQS_GT_=32 -DQS_CH_=256 -DQS_L2_HINTS=1" $EXTRA_VARIANTS; do
  rm -f qcp.o
  make EXTRA="$v" qcp.o libmdagpu.so > /dev/null 2>&1 || { echo "build failed: $v"; continue; }
  echo "== $v"
  (cd ../..; python profiles/microbench/bench_kernels.py qcp 2>&1 | grep "^qcp" | cut -c1-150)
done
rm -f qcp.o; make qcp.o libmdagpu.so > /dev/null 2>&1
