# usage: run_ab.sh variant...   (inside gpurun)
for v in "$@"; do
  if [ "$v" = "default" ]; then unset TGPU_LIB; else export TGPU_LIB=$PWD/tristan_mp_pu_master_densdecomp_b200/libtristan_gpu_$v.so; fi
  python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_$v.log 2>gpurun_out/bench_$v.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_$v.log').read().strip().splitlines()[-1]); print('$v', '%.4g'%d['value'], '%.2f'%d['ms_per_step'], {k:round(x,2) for k,x in d['roofline']['phase_ms'].items()})
"
done
