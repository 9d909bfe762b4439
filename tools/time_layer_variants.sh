#!/bin/bash
for v in "CNF_X=0" "CNF_GC_V2=1" "CNF_GC_TC=1" "CNF_PW_TC=0"; do
  echo -n "$v : "; env $v python tools/profile_layer.py 256 5 | tail -1
done
