run() { name=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29578 bench.py --gpus 4 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; }
run r2_bench_4gpu_driver --steps 20 --warmup 5 --no-steady
run r2_bench_4gpu_default --steps 60 --warmup 30 --no-steady
