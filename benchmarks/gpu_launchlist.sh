#!/bin/bash
# development helper: per-launch device times (ncu gpu__time_duration) of bench.py with the given options
tag=$1; skip=$2; cnt=$3; shift 3
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary "$@" > gpurun_out/plain_ll_$tag.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s $skip -c $cnt --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary "$@" > gpurun_out/ncu_ll_$tag.log 2>&1; echo ncu_rc=$?
cat gpurun_out/plain_ll_$tag.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('qps %.0f ms %.3f share %.3f launches/step %d' % (d['value'], d['ms_per_step'], d['roofline']['kernel_share_of_step'], d['gpu_launches']/d['steps']))"
