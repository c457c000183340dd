timeout 600 python -m pytest tests -m gpu -x -q -k "dithered" 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"rng_states|import_fast" --csv python tools/_dither_probe.py 2>&1 | grep -E "rng_states|import_fast" | awk -F'","' '{print substr($5,1,60), $(NF-2), $NF}' | tail -6
B2_FRAMES=256 timeout 300 python tools/quick_bench.py import 2>&1 | grep DITHER
