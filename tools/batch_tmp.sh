python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in 1 16 24 32; do
  echo "config4 [cr$v]"; FT_LIB=build/lib_cr$v.so python bench.py --workload config4 --no-cpu --steps 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['frac'])"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_p2p.py > gpurun_out/r2v_check_p2p.log 2>&1; tail -5 gpurun_out/r2v_check_p2p.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2v_bench2.json 2> gpurun_out/r2v_bench2.err; tail -3 gpurun_out/r2v_bench2.err; cat gpurun_out/r2v_bench2.json
