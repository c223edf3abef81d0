#!/bin/bash
# A/B runs of the headline bench with engine knobs (GPU box):  bash tools/ab_bench.sh <tag> "name:GLOBAL_OPTS:ENGINE_OPTS" ...
tag=$1; shift
mkdir -p gpurun_out
for spec in "$@"; do
  name=${spec%%:*}; rest=${spec#*:}; gopt=${rest%%:*}; eopt=${rest#*:}
  BENCH_GLOBAL_OPTIONS="$gopt" BENCH_ENGINE_OPTIONS="$eopt" python bench.py --steps 20 --warmup 3 --no-cpu-baseline \
      > gpurun_out/${tag}_ab_${name}.json 2> gpurun_out/${tag}_ab_${name}.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/${tag}_ab_${name}.json').read().strip().splitlines()[-1])
    print('${name}', 'value %.1f fps  e2e %.1f  prog %.1f us  frac %.3f  parity epe %.4f idx %.4f' % (d['value'], d['e2e']['value'], d['roofline']['us_per_launch'] or 0,
          d['roofline']['frac'] or 0, (d.get('parity') or {}).get('epe_mean_px', -1), (d.get('parity') or {}).get('index_agreement', -1)))
except Exception as ex:
    print('${name}', 'FAILED', ex, open('gpurun_out/${tag}_ab_${name}.err').read()[-400:])
PY
done
