#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x -k "upsample or decoder or 16bit or emit or config3 or config4 or rescaled" 2>&1 | tail -5 | cut -c1-300
python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
from go_webp_b200 import _abi
from go_webp_b200.device import Context
from go_webp_b200.pipeline import UpsamplePipeline
ctx = Context(0, 1)
for mode, name, bpp, wa in ((_abi.MODE_RGB_565, "RGB_565", 3.5, False), (_abi.MODE_RGBA_4444, "RGBA_4444", 3.5, False), (_abi.MODE_RGBA_4444, "RGBA_4444+A", 4.5, True), (_abi.MODE_RGBA, "RGBA", 5.5, False)):
    p = UpsamplePipeline(ctx, 128, 1920, 1080, mode, with_alpha=wa, host=False)
    p.fill_synthetic(0xB2000004)
    for _ in range(3): p.run_device()
    ctx.sync(); ctx.event_record(0)
    for _ in range(10): p.run_device()
    ctx.event_record(1); ctx.sync()
    ms = ctx.elapsed_ms(0, 1) / 10
    px = 128 * 1920 * 1080
    print(f"upsample 128x1080p {name:12s} {px / ms / 1e3:10.0f} Mpx/s {px * bpp / ms / 1e6:7.0f} GB/s {px * bpp / ms / 1e6 / 6552:6.3f}")
    p.close()
PY
