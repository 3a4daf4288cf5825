set -x
run() { name=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus 8 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; }
run r2c_bench_8gpu_driver --steps 20 --warmup 5 --no-steady
run r2c_bench_8gpu_default --steps 60 --warmup 30 --no-steady
