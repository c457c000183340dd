set -x
python bench.py > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err
tail -c 600 gpurun_out/bench_r1c.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1c_ref.json 2>> gpurun_out/bench_r1c.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/ncu_launches.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:import_fast -c 1 -o gpurun_out/prof_r1c_import python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/ncu_imp_full.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"upsample|color_index|flat" -c 12 -o gpurun_out/prof_r1c_others python tools/profile_kernels.py upsample vp8l > gpurun_out/ncu_others.log 2>&1
