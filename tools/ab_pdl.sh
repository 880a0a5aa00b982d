for v in "" "IE_PDL=0" "IE_PDL=0 IE_SPLITK=0" "IE_SPLITK=0" ""; do
  env $v python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline --sustain 0 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$v', round(d['value'],1), round(d['ms_per_step'],3), 'conv', round(d['roofline']['conv_ms_per_step'],3), d['clocks']['sm_mhz'], d['clocks'].get('power_w'))"
done
