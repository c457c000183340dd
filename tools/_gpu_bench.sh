# Round-end style GPU pass: tests, bench (both arms), ncu launch list, ncu full captures.  TAG names the outputs.
TAG=${TAG:-r1e}
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -3 gpurun_out/pytest_gpu_$TAG.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 400 gpurun_out/bench_$TAG.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2>> gpurun_out/bench_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/ncu_launches_$TAG.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:import_fast -c 1 -o gpurun_out/prof_${TAG}_import python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/ncu_imp_full_$TAG.log 2>&1
