for v in w1 w1r64 w2r64 w4 w8; do
  B200GS_LIB=$PWD/variants/lib_$v.so timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/sweep_$v.json 2>/dev/null
  python -c "
import json
d=json.load(open('gpurun_out/sweep_$v.json'))
s=d['roofline']['stages']
print('$v', round(d['value'],1), 'fwd', round(s['blend_fwd']['ms_per_launch_set'],2), 'bwd', round(s['blend_bwd']['ms_per_launch_set'],2))"
done
