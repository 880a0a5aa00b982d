#!/bin/bash
# One GPU box, one pass: the evidence a round commits under profiles/ (see profiles/README.md).
#   gpurun --timeout 1800 -- 'bash tools/round_evidence.sh r02f [notest] [nostep]'
# gpurun brings back at most 64 MiB of gpurun_out/: the ncu reports stay in /tmp on the box and are condensed there.
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
timeout 400 python bench.py --steps 20 --warmup 5 --breakdown > $O/bench_cfg2_$TAG.json 2> $O/bench_cfg2_$TAG.err || echo "bench cfg2 failed"
timeout 200 python bench.py --config cfg1 --steps 50 --warmup 10 --no-extra > $O/bench_cfg1_$TAG.json 2> $O/bench_cfg1_$TAG.err || echo "bench cfg1 failed"
for CFG in cfg3 cfg4; do
  timeout 300 python bench.py --config $CFG --steps 5 --warmup 3 --no-extra --sustain 0 > $O/bench_${CFG}_$TAG.json 2> $O/bench_${CFG}_$TAG.err || echo "bench $CFG failed"
done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_$TAG.json 2> $O/bench_reference_$TAG.err || echo "reference arm failed"
timeout 200 python tools/latency_small.py > $O/latency_small_$TAG.txt 2>&1 || echo "latency failed"
timeout 200 python tools/basis_kpn_bench.py > $O/basis_kpn_bench_$TAG.jsonl 2> $O/basis_kpn_bench_$TAG.err || echo "basis_kpn failed"
timeout 100 python tools/layer_times.py > $O/layer_times_simplemodel_$TAG.txt 2>&1
timeout 100 python tools/layer_times.py --model Basis_kpn --batch 256 --size 64 --burst 8 --bases 10 > $O/layer_times_basis_kpn_b10_$TAG.txt 2>&1
timeout 200 python tools/metric_sweep.py --batches 1 4 16 > $O/metric_sweep_$TAG.jsonl 2> $O/metric_sweep_$TAG.err || echo "sweep failed"
# launch list of the bench command (cold-cache, serialised: compare shares, not absolutes)
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --sustain 0 > /dev/null 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches_bench_cfg2_$TAG.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --sustain 0 > $O/ncu_launches_$TAG.log 2>&1
# one profiled step (forward + metrics), full sections
if [[ " $* " != *" nostep "* ]]; then
  timeout 120 python tools/profile_step.py > /dev/null 2>&1 && \
  timeout 900 ncu --set full --clock-control none --profile-from-start off -f -o /tmp/step_cfg2_$TAG \
    python tools/profile_step.py > $O/ncu_step_$TAG.log 2>&1 && \
  python tools/ncu_summary.py /tmp/step_cfg2_$TAG.ncu-rep > $O/ncu_step_cfg2_$TAG.txt && \
  python tools/conv_traffic.py $O/ncu_step_cfg2_$TAG.txt > $O/conv_traffic_$TAG.json
fi
du -sh $O; ls -la $O | tail -25
