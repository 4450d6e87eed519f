#!/bin/bash
# ncu full capture of the forward and inverse chunk kernels at bench shape (batch 1024). Usage: bash tools/gpu_prof.sh tag
TAG=${1:-prof}
OUT=gpurun_out
mkdir -p $OUT
ncu --set full --clock-control none --import-source on -k regex:"k_ring|k_chunk" -s 6 -c 2 -f -o $OUT/${TAG} \
    python bench.py --steps 2 --warmup 3 --batch 1024 --no-cpu-baseline > $OUT/${TAG}_run.log 2>&1
tail -3 $OUT/${TAG}_run.log
