ncu --set full --import-source on --clock-control none -k regex:"upsample_bulk|flat_kernel|color_index|predictor|cleanup_yuva|blend_yuva|import_fast|rng_states" -c 40 -o gpurun_out/prof_r1g_others python tools/profile_kernels.py dither upsample vp8l predictor tools > gpurun_out/ncu_others_r1g.log 2>&1
tail -2 gpurun_out/ncu_others_r1g.log
