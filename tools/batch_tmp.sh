python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r3s_bench_default.json 2> gpurun_out/r3s_bench_default.err; tail -2 gpurun_out/r3s_bench_default.err; python -c "
import json; d=json.load(open('gpurun_out/r3s_bench_default.json')); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], 'equal', d['roofline']['equal_work']['frac'], 'cpu', d['cpu_baseline']['value'], 'c5', d['config5_strong']['ms_per_frame'], d['config5_strong']['e2e_ms_per_frame'])"
for c in config3 config4; do python bench.py --workload $c --no-cpu --steps 4 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['workload'], d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'])"; done
