python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/r3e_bench_default.json 2> gpurun_out/r3e_bench_default.err; tail -2 gpurun_out/r3e_bench_default.err; cut -c1-400 gpurun_out/r3e_bench_default.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r3e_bench_ref.json 2> gpurun_out/r3e_bench_ref.err; cut -c1-300 gpurun_out/r3e_bench_ref.json
bash tools/gpu_prof.sh r3e
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
