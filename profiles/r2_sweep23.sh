#!/bin/bash
# ncu summaries of the generic kernel on the HCN_UT (one Davidson block of 27 vectors) and HNO3 LB6/LG7 shapes
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
timeout 600 ncu --set full --import-source on --clock-control none -k regex:sg4_term_kernel_generic -c 16 -o $O/r2s23_hno3 -f python profiles/gen_case.py hno3 1 > $O/r2s23_hno3.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:sg4_term_kernel_generic -c 16 -o $O/r2s23_hcn -f python profiles/gen_case.py hcn 27 > $O/r2s23_hcn.log 2>&1
ls -la $O/*.ncu-rep
