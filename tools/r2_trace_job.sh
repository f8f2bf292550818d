run() { tag=$1; shift; env "$@" TF_BENCH_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 5 --no-e2e --no-cpu --no-extras --no-gate > gpurun_out/r2_trace_$tag.json 2> gpurun_out/r2_trace_$tag.err; echo $tag rc=$?; }
run base A=1
run ctas4 NCCL_MAX_CTAS=4
run ctas2 NCCL_MAX_CTAS=2
run res16 TRACEFLY_RESERVE_SMS=16 NCCL_MAX_CTAS=8
run res0 TRACEFLY_RESERVE_SMS=0
