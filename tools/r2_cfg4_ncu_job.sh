set -x
# cfg4 after the phased list walk / compact list heads: plain run, then ncu --set full of the generation kernel at
# generation 6 (population still spread out) and at generation 15
B="python bench.py --no-e2e --no-cpu --no-gate --no-extras --workload cfg4"
$B --steps 20 --warmup 5 > gpurun_out/r2d_plain_cfg4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:duo_kernel -s 16 -c 1 -f -o gpurun_out/r2d_cfg4_gen15 $B --steps 20 --warmup 5 > gpurun_out/r2d_ncu_cfg4_gen15.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:duo_kernel -s 7 -c 1 -f -o gpurun_out/r2d_cfg4_gen6 $B --steps 20 --warmup 5 > gpurun_out/r2d_ncu_cfg4_gen6.log 2>&1
ls -la gpurun_out/*.ncu-rep
# cfg3 (default workload) at generation 15 with the same build, and the launch list of the default run
B3="python bench.py --no-e2e --no-cpu --no-gate --no-extras"
$B3 --steps 20 --warmup 5 > gpurun_out/r2d_plain_cfg3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:duo_kernel -s 16 -c 1 -f -o gpurun_out/r2d_cfg3_gen15 $B3 --steps 20 --warmup 5 > gpurun_out/r2d_ncu_cfg3_gen15.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2d_launches_default.csv $B3 --steps 20 --warmup 5 > gpurun_out/r2d_ncu_launches.log 2>&1
ls -la gpurun_out/*.ncu-rep
