# 4 slabs: parity against the single-GPU run (interior ranks have two neighbours), then the driver-window bench
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 tools/multigpu_check.py "$@" 2>&1 | grep "^\[f\|ok=\|Error\|error" | tail -3; }
run f64 small 25 0 0 1
run f64 small 25 0 0 0
run f32 small 25 2 0 1
run f64 small 25 0 1 0
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29578 bench.py --gpus 4 --steps 20 --warmup 5 --no-steady > gpurun_out/r2c_bench_4gpu_driver.json 2> gpurun_out/r2c_bench_4gpu_driver.err; echo "bench rc=$?"
