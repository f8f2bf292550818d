#!/bin/bash
# cfg4 generation time against the geometry of the moving-sphere index (bench.py reads TF_DYN_CELL / TF_DYN_BAND / TF_DYN_CAP)
cd "$(dirname "$0")/.."
GEOMS=${GEOMS:-8,16 8,8 8,4 6,16 6,8 12,8}
for g in $GEOMS; do
  c=${g%%,*}; b=${g##*,}
  TF_DYN_CELL=$c TF_DYN_BAND=$b python bench.py --workload cfg4 --no-e2e --no-cpu --no-extras --steps 20 --warmup 5 > gpurun_out/geom_${c}_${b}.json 2> gpurun_out/geom_${c}_${b}.err
  echo "cell $c band $b rc=$? $(python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/geom_${c}_${b}.json').read().strip().splitlines()[-1])
    r = d['roofline']
    print('kernel_ms', round(r['kernel_ms'], 4), 'q1', round(r['kernel_ms_first_quarter'], 4), 'q4', round(r['kernel_ms_last_quarter'], 4), 'gate', d.get('parity_gate', {}).get('ok'))
except Exception as ex:
    print('no line', ex)
PY
)"
done
