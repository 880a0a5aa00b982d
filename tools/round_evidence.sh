#!/bin/bash
# One GPU box, one pass: the evidence a round commits under profiles/ (see profiles/README.md).
#   gpurun --timeout 1500 -- 'bash tools/round_evidence.sh v5 [notest] [nostep]'
# gpurun brings back at most 64 MiB of gpurun_out/: the ncu reports are condensed on the box and large ones deleted.
# Every ncu pass runs only after the same command exited 0 without the profiler; numbers printed under ncu are never
# bench values.
set -u
TAG=${1:-vX}
O=gpurun_out
mkdir -p $O
if [[ " $* " != *" notest "* ]]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu_$TAG.log
  tail -3 $O/pytest_gpu_$TAG.log
fi
for CFG in cfg2 cfg3 cfg4; do
  timeout 300 python bench.py --config $CFG --steps 10 --warmup 3 --breakdown > $O/bench_${CFG}_$TAG.json 2> $O/bench_${CFG}_$TAG.err || echo "bench $CFG failed"
done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_$TAG.json 2> $O/bench_reference_$TAG.err || echo "reference arm failed"
timeout 200 python tools/metric_sweep.py --batches 1 2 4 8 16 --variants > $O/metric_sweep_$TAG.jsonl 2> $O/metric_sweep_$TAG.err || echo "sweep failed"
# launch list of the bench command (cold-cache, serialised: compare shares, not absolutes)
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches_bench_cfg2_$TAG.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_launches_$TAG.log 2>&1
# one profiled step (forward + metrics), full sections
if [[ " $* " != *" nostep "* ]]; then
  timeout 120 python tools/profile_step.py > /dev/null 2>&1 && \
  timeout 900 ncu --set full --clock-control none --profile-from-start off -f -o /tmp/step_cfg2_$TAG \
    python tools/profile_step.py > $O/ncu_step_$TAG.log 2>&1 && \
  python tools/ncu_summary.py /tmp/step_cfg2_$TAG.ncu-rep > $O/ncu_step_cfg2_$TAG.txt
fi
timeout 120 python tools/profile_metrics.py > /dev/null 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o $O/metrics_$TAG \
  python tools/profile_metrics.py > $O/ncu_metrics_$TAG.log 2>&1 && \
python tools/ncu_summary.py $O/metrics_$TAG.ncu-rep > $O/ncu_metrics_$TAG.txt
du -sh $O; ls -la $O | tail -15
