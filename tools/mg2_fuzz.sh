run() { FUZZ_SEED=$1 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/multigpu_check.py $2 small 20 $3 $4 $5 2>&1 | grep "^\[f\|Error\|error\|Traceback" | tail -2 | sed "s/^/seed $1: /"; }
for s in 1 2 3 4 5; do
  run $s f64 0 0 0
  run $s f32 2 0 1
  run $s f64 2 1 1
done
