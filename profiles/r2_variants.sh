#!/bin/bash
# the whole -m gpu suite under every kernel-selection switch (each variant is a different code path of the library)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
: > $O/r2_variants.txt
for v in "EVR_X=0" "EVR_SG4_V2=1" "EVR_SG4_V2=1 EVR_SG4_V2_THREADS=512" "EVR_SG4_ISO=0" "EVR_SG4_ISO=2" "EVR_SG4_FORCE_GENERIC=1" "EVR_SG4_BLOCK_ORDER=1" "EVR_SG4_BATCH=0" "EVR_SG4_DYNAMIC=1" "EVR_SG4_GRAPH=1" "EVR_SG4_DMMA=1" "EVR_SG4_DEBUG=768" "EVR_SG4_ORDER_ROUNDS=1" "EVR_SG4_BCAP=700" "EVR_SG4_DETERMINISTIC=1"; do
  echo "## $v" >> $O/r2_variants.txt
  env $v timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2 >> $O/r2_variants.txt
done
cat $O/r2_variants.txt
