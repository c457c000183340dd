timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r01h_bench.json 2> gpurun_out/r01h_bench.err; echo bench rc=$?
B2_FRAMES=128 timeout 600 python tools/quick_bench.py dequant filters rescale convert tools > gpurun_out/r01h_quick_next_rows.txt 2>&1; echo quick rc=$?
