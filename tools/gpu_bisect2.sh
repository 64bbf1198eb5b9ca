#!/bin/bash
run() { echo "== $*"; env QUICK=1 "$@" python tools/debug_batch.py 80 2 2>&1 | grep "QUICK\|determin\|Error" ; }
run A=1
run SDB200_COLSTATS=0
run SDB200_TC_KERNEL=single
run SDB200_TC_PLANS=0
run SDB200_GN=split
run SDB200_PDL=0
