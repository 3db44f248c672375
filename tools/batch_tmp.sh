python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/kexp.py "" "FT_LIB=build/lib_head.so" > gpurun_out/r3b_kexp.txt 2>&1; cut -c1-150 gpurun_out/r3b_kexp.txt
