#!/bin/bash
# times the fused grouped-conv kernel with phases switched off (CNF_OCT_DBG bits) -- results are wrong by design
for d in 0 1 2 4 8 16 3 7 15 31; do
  echo -n "dbg=$d: "
  CNF_OCT_DBG=$d ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gconv_oct -c 2 python tools/profile_layer.py 256 1 2>/dev/null | grep duration | tail -1
done
