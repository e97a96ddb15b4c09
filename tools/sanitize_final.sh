#!/bin/bash
# memcheck / initcheck of the final round-2 kernels (default path, chained-smoother path, march + tiny kernels forced on)
out=gpurun_out/san4; mkdir -p $out
CS=/usr/local/cuda/bin/compute-sanitizer
run() { name=$1; tool=$2; shift 2; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  echo "=== $name: compute-sanitizer --tool $tool ${envs[*]} python tools/sanitize_run.py $*" > $out/san_$name.txt
  env RLFC_NO_GRAPH=1 "${envs[@]}" timeout 600 $CS --tool $tool --print-limit 20 python tools/sanitize_run.py "$@" >> $out/san_$name.txt 2>&1
  echo "exit code $?" >> $out/san_$name.txt
  echo "$name: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $out/san_$name.txt | tail -1)"; }
run memcheck_default memcheck RLFC_RESID=march --
run initcheck_default initcheck RLFC_RESID=march --
run memcheck_chain memcheck RLFC_SMOOTHER=chain --
run racecheck_default racecheck RLFC_RESID=march -- --small
