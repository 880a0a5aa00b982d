O=gpurun_out; T=r02h
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$T.log 2>&1; tail -2 $O/pytest_gpu_$T.log
timeout 400 python bench.py --steps 20 --warmup 5 --breakdown > $O/bench_cfg2_$T.json 2> $O/bench_cfg2_$T.err
timeout 200 python bench.py --config cfg1 --steps 50 --warmup 10 --no-extra > $O/bench_cfg1_$T.json 2> $O/bench_cfg1_$T.err
timeout 200 python tools/basis_kpn_bench.py > $O/basis_kpn_bench_$T.jsonl 2> $O/basis_kpn_bench_$T.err
timeout 100 python tools/layer_times.py > $O/layer_times_simplemodel_$T.txt 2>&1
timeout 100 python tools/layer_times.py --model Basis_kpn --batch 256 --size 64 --burst 8 --bases 10 > $O/layer_times_basis_kpn_b10_$T.txt 2>&1
timeout 200 python tools/latency_small.py > $O/latency_small_$T.txt 2>&1
python -c "
import json
for c in ['cfg2','cfg1']:
    d=json.loads(open('$O/bench_%s_$T.json'%c).read().strip().splitlines()[-1]); print(c, round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'roof', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'], d.get('sustained',{}).get('value'))
"
cut -c1-110 $O/basis_kpn_bench_$T.jsonl; grep graph $O/latency_small_$T.txt | head -3
