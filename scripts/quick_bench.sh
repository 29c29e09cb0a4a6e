# usage (inside gpurun): scripts/quick_bench.sh tag [extra bench args]   -> prints value, ms per lap and the phase timers
tag=$1; shift
python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu "$@" > gpurun_out/bench_$tag.log 2>gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$tag.log').read().strip().splitlines()[-1])
print('$tag', '%.4g'%d['value'], '%.2f ms'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], {k:round(x,2) for k,x in d['roofline']['phase_ms'].items()})
PY
