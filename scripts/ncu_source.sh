#!/bin/bash
# Source-level stall samples of one launch of the step kernel: scripts/ncu_source.sh TAG [ENV=VAL ...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=$1; shift
env "$@" SJ_NO_GRAPH=1 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section MemoryWorkloadAnalysis --import-source on \
    --clock-control none --cache-control none -k regex:"tma" -s 14 -c 1 -f -o /tmp/ncu_src_$tag python scripts/prof_steps.py 30 > gpurun_out/ncu_src_$tag.log 2>&1
cp /tmp/ncu_src_$tag.ncu-rep gpurun_out/ncu_src_$tag.ncu-rep
ls -la gpurun_out/ncu_src_$tag.ncu-rep
