"""Build an instrumented / experimental variant of librlfc.so next to the shipped one (never loaded by default:
select it with RLFC_LIBRARY=<path>).  usage: build_variant.py <name> [-DMACRO[=v] ...]"""
import subprocess
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from rlfluidcontrol_b200 import build as B

name, defs = sys.argv[1], sys.argv[2:]
out = B.PKG / f"librlfc_{name}.so"
cmd = [B.nvcc_path(), *B.NVCC_FLAGS, *defs, "-o", str(out), *[str(B.CSRC / s) for s in B.SOURCES]]
res = subprocess.run(cmd, capture_output=True, text=True)
if res.returncode != 0:
    raise SystemExit(res.stdout + res.stderr)
print(out)
