#!/bin/bash
# End-of-round N = 8 lines of the two configs whose named workload is training (config 4) / the pre-training sweep (config 5),
# after the tensor-core grouped data gradients.  usage: bash tools/r02_multigpu_train.sh   (8 GPUs)
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out
P=29610
for c in 4 5; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 8 --config $c --steps 5 --warmup 3 --quick > $O/r02y_bench_line_cfg${c}_n8.json 2> $O/r02y_bench_line_cfg${c}_n8.err
  echo "cfg$c rc=$?"; P=$((P+1))
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 8 --steps 20 --warmup 3 --quick > $O/r02y_bench_line_cfg2_n8.json 2> $O/r02y_bench_line_cfg2_n8.err
echo "cfg2 rc=$?"
