#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both arms), ncu launch list and one full capture.
# Usage (from the repo root, under gpurun): bash tools/gpu_check.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee $OUT/${TAG}_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/${TAG}_smoke.txt
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $OUT/${TAG}_clocks.csv &
SMI=$!
python bench.py --steps 20 --warmup 5 2> $OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench.json
kill $SMI
python bench.py --impl reference --steps 5 --warmup 1 | tee $OUT/${TAG}_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --batch 1024 --no-cpu-baseline > $OUT/${TAG}_ncu_launch_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_ring|k_chunk' -s 6 -c 2 -f -o $OUT/${TAG}_prof \
    python bench.py --steps 2 --warmup 3 --batch 1024 --no-cpu-baseline > $OUT/${TAG}_ncu_full_run.log 2>&1
ls -la $OUT | tail -12
