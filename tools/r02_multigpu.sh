#!/bin/bash
# Round-2 multi-GPU bench lines (one box, torchrun, one rank per GPU).  usage: bash tools/r02_multigpu.sh 8 | 4
# N = 8: config 2 (eval + sampling + training), config 3 (eval + sampling), config 4 (training is its named workload),
#        config 5 (noise pre-training sweep point N = 8);  N = 4: the N = 4 and N = 2 points of the config-5 sweep and config 4.
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out
P=29510
run() {   # run <nproc> <outfile> <bench args...>
  local n=$1 out=$2; shift 2
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P bench.py --gpus $n "$@" > $O/$out 2> $O/${out%.json}.err
  echo "$out rc=$? $(python -c "import json,sys; d=json.load(open('$O/$out')); t=d['config'].get('train_step') or {}; print(round(d['value']), 'img/s', round(d['ms_per_step'],2), 'ms; train', t.get('images_per_s'), t.get('ms_per_step'))" 2>&1 | tail -1)"
  P=$((P+1))
}
python -m pytest tests/test_gpu_multi_device.py -q -m gpu 2>&1 | tail -2 > $O/r02o_two_device_test.log; cat $O/r02o_two_device_test.log
if [ "$1" = "8" ]; then
  run 8 r02o_bench_line_cfg2_n8.json --steps 20 --warmup 3 --quick
  run 8 r02o_bench_line_cfg3_n8.json --config 3 --steps 10 --warmup 3 --quick
  run 8 r02o_bench_line_cfg4_n8.json --config 4 --steps 5 --warmup 3 --quick
  run 8 r02o_bench_line_cfg5_n8.json --config 5 --steps 5 --warmup 3 --quick
else
  run 4 r02o_bench_line_cfg5_n4.json --config 5 --steps 5 --warmup 3 --quick
  run 2 r02o_bench_line_cfg5_n2.json --config 5 --steps 5 --warmup 3 --quick
  run 4 r02o_bench_line_cfg4_n4.json --config 4 --steps 5 --warmup 3 --quick
  run 2 r02o_bench_line_cfg3_n2.json --config 3 --steps 10 --warmup 3 --quick
fi
