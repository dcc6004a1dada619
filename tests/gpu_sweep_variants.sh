# ad-hoc: bench several builds of the same C-ABI (variants/lib_<name>.so) through B200GS_LIB
for v in "$@"; do
  B200GS_LIB=$PWD/variants/lib_$v.so timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/sweep_$v.json 2>/dev/null
  python -c "
import json
d=json.load(open('gpurun_out/sweep_$v.json'))
s=d['roofline']['stages']
print('$v', round(d['value'],1), ' '.join(k+'='+str(round(x['ms_per_launch_set'],2)) for k,x in s.items()))"
done
