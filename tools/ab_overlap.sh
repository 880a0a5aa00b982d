# A-B of the side-stream basis branch (Engine option overlap_branches), interleaved: latency, cfg1, cfg2, Basis_kpn
set -x
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_ref_golden.py tests/test_gpu_eval.py -m gpu -x -q 2>&1 | tail -3
for rep in 1 2; do
for v in 0 1; do
  echo "=== IE_OVERLAP=$v"
  IE_OVERLAP=$v python tools/latency_small.py 2>&1 | grep -v "^$" | grep "graph\|eager: 32x\|eager: 128x"
  IE_OVERLAP=$v python tools/basis_kpn_bench.py --bases 10 2>&1 | tail -1 | cut -c1-120
  IE_OVERLAP=$v python bench.py --config cfg1 --steps 50 --warmup 10 --no-extra --sustain 0 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg1', d['value'], d['ms_per_step'])"
  IE_OVERLAP=$v python bench.py --steps 10 --warmup 3 --no-extra --sustain 0 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg2', d['value'], d['ms_per_step'], d['e2e']['value'])"
done; done
