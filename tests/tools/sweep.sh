#!/bin/bash
# Dev tool (GPU box): sweep kernel variants / scheduling knobs with kbench. usage: sweep.sh out.log "ENV1=.. ENV2=.." ...
out=$1; shift
: > $out
for cfg in "$@"; do
  echo "=== $cfg" >> $out
  env $cfg timeout 120 python tests/tools/kbench.py C2 C3 2>&1 | grep -E "ndiff [1-9]|isect=|Error|error" >> $out
done
cat $out
