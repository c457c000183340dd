B2_FRAMES=256 timeout 300 python tools/quick_bench.py import 2>&1 | tail -6
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"rng_states|import_fast" --launch-skip 60 --launch-count 6 --csv python tools/quick_bench.py import 2>&1 | grep -E "rng_states|import_fast" | awk -F'","' '{print $5, $NF}' | cut -c1-160 | tail -6
