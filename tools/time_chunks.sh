#!/bin/bash
for c in 0 128 64 32 16; do
  echo -n "chunk=$c : "; CNF_BATCH_CHUNK=$c python bench.py --steps 5 --warmup 3 --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), 'img/s', round(d['ms_per_step'],2),'ms/step')"
done
