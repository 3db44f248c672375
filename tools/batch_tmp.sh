for v in "" r8m7 r16m7 r24m7 r16m6 r8m6; do
  echo "config3 [$v]"; FT_LIB=${v:+build/lib_$v.so} python bench.py --workload config3 --no-cpu --steps 4 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['frac'])"
done
