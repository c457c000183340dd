set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r01h_bench.json 2> gpurun_out/r01h_bench.err
B2_FRAMES=64 timeout 600 python tools/quick_bench.py dequant filters rescale > gpurun_out/r01h_quick_next_rows.txt 2>&1
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
B2_FRAMES=64 timeout 600 ncu --metrics $M --clock-control none -k regex:'dq_|rescale_kernel|gradient' -c 60 --csv --log-file gpurun_out/r01h_ncu_next_rows.csv python tools/quick_bench.py dequant filters rescale > gpurun_out/r01h_ncu.log 2>&1
tail -2 gpurun_out/r01h_ncu.log
