set -x
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py tests/test_ref_golden.py -m gpu -x -q 2>&1 | tail -3
for rep in 1 2; do
for v in "IE_SPLITK_WIDE=0" "IE_SPLITK_MAX=8" "IE_SPLITK_MAX=16" "IE_SPLITK_MAX=36"; do
  echo "=== $v"
  env $v python tools/latency_small.py 2>&1 | grep -v "^$" | head -12
  env $v python tools/basis_kpn_bench.py --bases 10 50 2>&1 | tail -2 | cut -c1-400
  env $v python bench.py --steps 10 --warmup 3 --no-extra --sustain 0 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg2', d['value'], d['ms_per_step'])"
done; done
