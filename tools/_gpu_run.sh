for nf in 256 96; do
for f in "" go_webp_b200/lib/variants/*.so; do
  n=$(basename "$f" .so); n=${n#libb2webp_}
  B2_FRAMES=$nf B2_LIB=$f python tools/quick_bench.py import 2>&1 | tail -4 | awk -v n="${n:-default}-$nf" '{printf "%s %s=%s ", (NR==1?n:""), $3, $(NF-3)} END{print ""}'
done; done > gpurun_out/import_grid6.log 2>&1
cat gpurun_out/import_grid6.log
