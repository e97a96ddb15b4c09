#!/bin/sh
# Regenerates rlfluidcontrol_b200/data/init_state.bdimb from the reference's only fixture (needs /root/reference; run in the
# build container, not on the GPU box).  The text checkpoint (BDIM.write format, BDIM.pde:226-237) is parsed with
# strtof by the oracle's reader and stored as raw little-endian float32: "RLFCBDIM", int32 n, int32 m, float t,
# float dt, ux[n*m], uy[n*m], p[n*m].
set -e
cd "$(dirname "$0")/../.."
make -C oracle -s
oracle/_build/oracle_cli convert /root/reference/clientLilypad/saved/init/init.bdim rlfluidcontrol_b200/data/init_state.bdimb 386 194
