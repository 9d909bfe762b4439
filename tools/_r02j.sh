cd $GRAFT_REPO_ROOT
python tools/profile_layer.py 256 2 2 small > gpurun_out/r02j_small_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fused_coupling -s 2 -c 1 -f -o gpurun_out/r02j_small_full python tools/profile_layer.py 256 2 2 small > gpurun_out/r02j_small_ncu.log 2>&1
for c in 4 5; do
python bench.py --config $c --steps 1 --warmup 1 --no-train --no-cpu-baseline > gpurun_out/r02j_cfg${c}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02j_launches_cfg$c.csv python bench.py --config $c --steps 1 --warmup 1 --no-train --no-cpu-baseline > gpurun_out/r02j_cfg${c}_ncu.log 2>&1
done
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
