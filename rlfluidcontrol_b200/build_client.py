"""Build rlfc_client: the C re-implementation of clientLilypad/clientCFD.pde on top of librlfc.so."""
from __future__ import annotations

import subprocess
from pathlib import Path

from . import build as product_build

PKG = Path(__file__).resolve().parent
SRC = PKG / "csrc" / "rlfc_client.c"
BIN = PKG / "rlfc_client"


def build(force: bool = False) -> Path:
    lib = product_build.build()
    if not force and BIN.exists() and BIN.stat().st_mtime >= max(SRC.stat().st_mtime, lib.stat().st_mtime):
        return BIN
    cmd = ["gcc", "-O2", "-std=gnu99", "-Wall", "-o", str(BIN), str(SRC), f"-L{PKG}", "-lrlfc", "-Wl,-rpath,$ORIGIN"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("gcc failed:\n" + res.stdout + res.stderr)
    return BIN


if __name__ == "__main__":
    print(build(force=True))
