set -x
B="python bench.py --no-e2e --no-cpu --no-gate --no-extras"
$B --steps 20 --warmup 5 > gpurun_out/r2_plain_cfg3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:duo_kernel -s 16 -c 1 -f -o gpurun_out/r2_cfg3_gen15 $B --steps 20 --warmup 5 > gpurun_out/r2_ncu_cfg3_gen15.log 2>&1
$B --steps 300 --warmup 5 > gpurun_out/r2_plain_cfg3_long.log 2>&1 && \
ncu --set full --clock-control none -k regex:duo_kernel -s 251 -c 1 -f -o gpurun_out/r2_cfg3_gen250 $B --steps 300 --warmup 5 > gpurun_out/r2_ncu_cfg3_gen250.log 2>&1
$B --workload cfg4 --steps 20 --warmup 5 > gpurun_out/r2_plain_cfg4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:duo_kernel -s 16 -c 1 -f -o gpurun_out/r2_cfg4_gen15 $B --workload cfg4 --steps 20 --warmup 5 > gpurun_out/r2_ncu_cfg4_gen15.log 2>&1
$B --workload cfg2 --steps 20 --warmup 5 > gpurun_out/r2_plain_cfg2.log 2>&1 && \
ncu --set full --clock-control none -k regex:stream_kernel -s 16 -c 1 -f -o gpurun_out/r2_cfg2_gen15 $B --workload cfg2 --steps 20 --warmup 5 > gpurun_out/r2_ncu_cfg2_gen15.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_default.csv $B --steps 20 --warmup 5 > gpurun_out/r2_ncu_launches.log 2>&1
ls -la gpurun_out/*.ncu-rep
