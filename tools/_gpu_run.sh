timeout 600 python tools/_e2e_probe.py 2>&1 | tail -14
