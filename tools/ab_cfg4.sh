#!/bin/bash
# A/B of kernel variants on cfg4 (and cfg3): every gpurun_ab/<name>/libtracefly_b200.so (build.py with TF_BUILD_DIR /
# TF_NVCC_EXTRA) runs bench.py's cfg4 workload with the parity gate, then one generation kernel under ncu for its
# instruction count.  usage: tools/ab_cfg4.sh [variant ...]   (default: all)   WL="cfg4 cfg3" to choose workloads
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
VARS=${@:-$(ls gpurun_ab)}
for v in $VARS; do
  lib=$PWD/gpurun_ab/$v/libtracefly_b200.so
  [ "$v" = product ] && lib=$PWD/tracefly-ga-uav-pathplanning_b200/_build/libtracefly_b200.so
  [ -f "$lib" ] || continue
  for wl in ${WL:-cfg4}; do
    TRACEFLY_B200_LIB=$lib python bench.py --workload $wl --no-e2e --no-cpu --no-extras --steps 20 --warmup 5 > gpurun_out/ab_${v}_${wl}.json 2> gpurun_out/ab_${v}_${wl}.err
    echo "$v $wl rc=$? $(python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/ab_${v}_${wl}.json').read().strip().splitlines()[-1])
    r = d['roofline']
    print('kernel_ms', round(r['kernel_ms'], 4), 'q1', round(r['kernel_ms_first_quarter'], 4), 'q4', round(r['kernel_ms_last_quarter'], 4), 'step', round(d['ms_per_step'], 4), 'gate', d.get('parity_gate', {}).get('ok'))
except Exception as ex:
    print('no line', ex)
PY
)"
    if [ -n "$NCU" ]; then
      TRACEFLY_B200_LIB=$lib ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,launch__registers_per_thread --clock-control none -k regex:duo_kernel -s 16 -c 1 --csv python bench.py --workload $wl --no-e2e --no-cpu --no-gate --no-extras --steps 20 --warmup 5 2>/dev/null | grep -E "inst_executed|time_duration|registers" | awk -F'","' '{print "   ncu", $(NF-2), $NF}' | tr -d '"'
    fi
  done
done
