#!/bin/bash
# End-of-round refresh after the warp-specialised octet kernel: bench line (full), launch list of one step, ncu --set full of the
# kernels of a 28x28x64 layer (exported to CSV on the box).  Outputs: gpurun_out/r02z_*
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out
python bench.py --steps 20 --warmup 3 > $O/r02z_bench_line.json 2> $O/r02z_bench.err; echo "bench rc=$?"
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-train --quick > $O/r02z_step_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r02z_launches_bench_step.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-train --quick > $O/r02z_step_ncu.log 2>&1
python tools/profile_layer.py 256 1 > $O/r02z_layer_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -s 11 -c 11 -f -o /tmp/r02z_layer_full python tools/profile_layer.py 256 1 > $O/r02z_layer_ncu.log 2>&1
python tools/ncu_export.py /tmp/r02z_layer_full.ncu-rep $O/r02z_layer28x28x64_ncu_full gconv_oct pw_tc3
grep -h "layer fwd" $O/r02z_layer_plain.log
