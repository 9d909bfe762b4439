#!/bin/bash
# End-of-round refresh (TMA producer + elected MMA issuer in the 1x1 kernel, two-stream step): full bench line, launch list of one
# step (streams = 1: ncu serialises launches anyway), ncu --set full of the kernels of a 28x28x64 layer (exported to CSV on the
# box), quick lines of configs 3 / 4 / 5.  Outputs: gpurun_out/r02f_*
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out
python bench.py --steps 20 --warmup 3 > $O/r02f_bench_line.json 2> $O/r02f_bench.err; echo "bench rc=$?"
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-train --quick --streams 1 > $O/r02f_step_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r02f_launches_bench_step.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-train --quick --streams 1 > $O/r02f_step_ncu.log 2>&1
python tools/profile_layer.py 256 1 > $O/r02f_layer_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -s 11 -c 11 -f -o /tmp/r02f_layer_full python tools/profile_layer.py 256 1 > $O/r02f_layer_ncu.log 2>&1
python tools/ncu_export.py /tmp/r02f_layer_full.ncu-rep $O/r02f_layer28x28x64_ncu_full gconv_oct pw_tc3
grep -h "layer fwd" $O/r02f_layer_plain.log
for c in 3 4 5; do
  python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --quick > $O/r02f_bench_line_cfg$c.json 2> $O/r02f_bench_cfg$c.err; echo "cfg$c rc=$?"
done
