"""A/B of the latency path on BASELINE configs 1-3 (one GPU): us per iteration of complete runs with the
environment given on the command line (e.g. BLK_PDL=0 vs default)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import blk_lanczos_b200 as B
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("BLK_")}, "small_configs": bench.small_configs(B, np)}))
