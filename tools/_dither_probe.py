import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from go_webp_b200.device import Context
from go_webp_b200.pipeline import ImportPipeline
ctx = Context(0, 1)
p = ImportPipeline(ctx, 256, 1920, 1080, host=False)
p.fill_synthetic(0xB2000001, 0, 2)
for _ in range(3):
    p.run_device(0, 0.5)
ctx.sync()
