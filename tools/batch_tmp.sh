python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/kexp.py "" "FT_LIB=build/lib_head.so" > gpurun_out/r2u_kexp.txt 2>&1; cut -c1-150 gpurun_out/r2u_kexp.txt
for v in "" nosplit; do
  echo "config3 [$v]"; FT_SPEC_NO_SPLIT=${v:+1} python bench.py --workload config3 --no-cpu --steps 4 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['frac'])"
done
for v in "" cr4 cr8 cr16; do
  echo "config4 [$v]"; FT_LIB=${v:+build/lib_$v.so} python bench.py --workload config4 --no-cpu --steps 2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['frac'])"
done
