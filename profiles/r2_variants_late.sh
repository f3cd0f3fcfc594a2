#!/bin/bash
# the -m gpu suite of the LAST build under the switches whose code paths changed late in round 2 (generic remainder,
# global-buffer classes, pipelined host blocks, graph capture of the remainder launches)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
: > $O/r2_variants_late.txt
for v in "EVR_SG4_GRAPH=1" "EVR_SG4_DETERMINISTIC=1" "EVR_SG4_FORCE_GENERIC=1" "EVR_SG4_ISO=0" "EVR_SG4_MIXED=0" "EVR_SG4_PIPELINE=0" "EVR_SG4_BLOCK_ORDER=1" "EVR_SG4_V2=1"; do
  echo "## $v" >> $O/r2_variants_late.txt
  env $v timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4 | cut -c1-300 >> $O/r2_variants_late.txt
done
cat $O/r2_variants_late.txt
