set -x
nvidia-smi topo -m > gpurun_out/r2_topo_8gpu.txt 2>&1
lscpu | grep -i "numa\|socket\|model name\|^CPU(s)" >> gpurun_out/r2_topo_8gpu.txt 2>&1
cat /sys/fs/cgroup/cpuset.cpus.effective /sys/fs/cgroup/cpuset.mems.effective >> gpurun_out/r2_topo_8gpu.txt 2>&1
run() { name=$1; shift; timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus 8 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; tail -c 600 gpurun_out/$name.json; }
run r2_bench_8gpu_driver --steps 20 --warmup 5 --no-steady
run r2_bench_8gpu_default --steps 60 --warmup 30 --no-steady
run r2_bench_8gpu_xslabs --steps 60 --warmup 30 --slab-axis 0 --no-steady
run r2_bench_50M_strong_8gpu --particles 50M --scaling strong --steps 30 --warmup 30 --no-steady
