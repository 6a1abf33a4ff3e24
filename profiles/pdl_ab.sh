BENDY_PDL=2 timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for L in 0 1 2; do
  BENDY_PDL=$L timeout 100 python bench.py --no-scaling-ref --no-cpu-baseline --no-e2e > gpurun_out/bench_pdl$L.json 2> gpurun_out/bench_pdl$L.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_pdl$L.json').read().strip().splitlines()[-1])
print('PDL=$L', d['value'], d['ms_per_step'], [round(x,3) for x in d['ms_per_step_series'][:4]], d['roofline']['substep']['frac'])
PY
done
