#!/bin/bash
# diagnostic: several multigpu_check variants, summary lines only
run() { # p2p prec name steps axis host
  DFSPH_B200_P2P=$1 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/multigpu_check.py $2 $3 $4 $5 $6 2>&1 | grep -A1 "^\[f" | sed "s/^/p2p=$1 /"
}
run 1 f32 small 5 2 0
run 1 f32 small 12 2 0
run 0 f32 small 25 2 0
run 1 f32 small 25 0 0
