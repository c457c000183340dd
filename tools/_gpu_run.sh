#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x -k "palette" 2>&1 | tail -3 | cut -c1-300
python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
from go_webp_b200.device import Context
from go_webp_b200.pipeline import PalettePipeline
ctx = Context(0, 1)
for w in (4096, 4095, 4094):
    for psize, bpp in ((256, 8.0),):
        p = PalettePipeline(ctx, 16, w, 4096, psize)
        p.fill_synthetic(0xB2000003)
        for _ in range(3): p.run_device()
        ctx.sync(); ctx.event_record(0)
        for _ in range(10): p.run_device()
        ctx.event_record(1); ctx.sync()
        ms = ctx.elapsed_ms(0, 1) / 10
        px = 16 * w * 4096
        print(f"apply_palette {psize:3d} colours 16x{w}x4096 {px / ms / 1e3:10.0f} Mpx/s {px * bpp / ms / 1e6 / 6552:6.3f}")
        p.close()
PY
