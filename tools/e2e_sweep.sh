#!/bin/bash
# sweep of the host-buffer pipeline parameters (chunks in flight x chunk size) for the e2e figure
for d in 2 3 4 6; do for mib in 16 32 64 128; do
  echo -n "depth $d chunk ${mib}MiB: "
  NTT_B200_PIPE_DEPTH=$d NTT_B200_PIPE_MIB=$mib python bench.py --steps 3 --warmup 3 --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.0f NTT/s  %.2f ms/step' % (d['e2e']['value'], d['e2e']['ms_per_step']))"
done; done
