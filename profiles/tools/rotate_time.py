"""Times RotateHoisted (8 rotations per decomposition, 32 ciphertexts) of one library build: LATTIGPU_LIB=... python rotate_time.py"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu-baseline", "--no-e2e", "--no-parity-check", "--steps", "6"],
                              stderr=subprocess.DEVNULL, text=True)
d = json.loads(out[out.index("{"):])
print(json.dumps({"lib": os.environ.get("LATTIGPU_LIB", "default").split("/")[-1], "hoisted": d["rotate"]["rotate_hoisted_rotations_per_s"],
                  "rotate": d["rotate"]["rotate_columns_ops_per_s"], "value": d["value"]}), flush=True)
