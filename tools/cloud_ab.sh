#!/bin/bash
# usage: bash tools/cloud_ab.sh lib1.so lib2.so ...   -- config-4 frame time under each library build ("" = the in-tree build)
for L in "$@"; do
  if [ -n "$L" ]; then export FT_LIB=$PWD/$L; else unset FT_LIB; fi
  R=$(timeout 8 python bench.py --workload config4 --steps 3 --warmup 2 --no-cpu --no-config5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'])")
  echo "${L:-default} $R"
done
