#!/bin/bash
# Round-2 single-GPU evidence: GPU tests, bench lines of every config, launch list of one config-2 step, ncu --set full of the
# kernels of a 28x28x64 layer, of the activation-resident kernel and of the tensor-core grouped conv.  Every ncu command runs
# after the same command exited 0 without ncu.  Outputs: gpurun_out/r02w_*
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02w_gpu_tests.log 2>&1; echo "tests rc=$?" >> $O/r02w_gpu_tests.log; tail -3 $O/r02w_gpu_tests.log
python bench.py --steps 20 --warmup 3 > $O/r02w_bench_line.json 2> $O/r02w_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $O/r02w_reference_arm_line.json 2>> $O/r02w_bench.err
for c in 3 4 4-heavy 5 5-heavy; do
  timeout 400 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > $O/r02w_bench_line_cfg$c.json 2> $O/r02w_bench_cfg$c.err; echo "cfg $c rc=$?"
done
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-train --quick > $O/r02w_step_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r02w_launches_bench_step.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-train --quick > $O/r02w_step_ncu.log 2>&1
python tools/profile_layer.py 256 1 > $O/r02w_layer_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -s 11 -c 11 -f -o /tmp/r02w_layer_full python tools/profile_layer.py 256 1 > $O/r02w_layer_ncu.log 2>&1
python tools/ncu_export.py /tmp/r02w_layer_full.ncu-rep $O/r02w_layer28x28x64_ncu_full gconv_oct pw_tc3
python tools/profile_layer.py 256 2 2 small > $O/r02w_small_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fused_coupling -s 2 -c 1 -f -o /tmp/r02w_small_full python tools/profile_layer.py 256 2 2 small > $O/r02w_small_ncu.log 2>&1
python tools/ncu_export.py /tmp/r02w_small_full.ncu-rep $O/r02w_fused_coupling_14x14x32_ncu_full fused_coupling
python tools/profile_layer.py 16 2 2 cfg5 > $O/r02w_cfg5layer_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gconv_tc -s 4 -c 2 -f -o /tmp/r02w_gconv_tc_full python tools/profile_layer.py 16 2 2 cfg5 > $O/r02w_cfg5layer_ncu.log 2>&1
python tools/ncu_export.py /tmp/r02w_gconv_tc_full.ncu-rep $O/r02w_gconv_tc_128x128x64_ncu_full gconv_tc
python tools/profile_train.py 256 1 once > $O/r02w_train_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -s 4 -c 2 -f -o /tmp/r02w_wgrad_tc_full python tools/profile_train.py 256 1 once > $O/r02w_train_ncu.log 2>&1
python tools/ncu_export.py /tmp/r02w_wgrad_tc_full.ncu-rep $O/r02w_wgrad_tc_ncu_full wgrad_tc_kernel
grep -h "layer fwd" $O/r02w_layer_plain.log $O/r02w_small_plain.log $O/r02w_cfg5layer_plain.log
