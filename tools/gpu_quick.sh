#!/bin/bash
# quick GPU check: parity tests then a short bench line (kernel ms, fractions)
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 --no-cpu-baseline | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.3fM NTT/s  fwd %.3f ms (%.1f%% hbm)  inv %.3f ms (%.1f%%)  e2e %.0f  clocks %s' % (d['value']/1e6, r['kernel_ms'], 100*r['frac'], r['inverse']['kernel_ms'], 100*r['inverse']['frac'], d['e2e']['value'], d['clocks']))"
