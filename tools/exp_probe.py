import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hsa_b200 import api, build
build.build_native()
for fp in (23 << 20, 46 << 20, 1550 << 20, 3100 << 20, 8 << 30):
    print(fp >> 20, "MB:", round(api.random_sector_probe(0, fp, 64), 1), "GB/s", flush=True)
