#!/bin/bash
# e2e against how the host widens the fp32 bands (store kind, worker count).
out=gpurun_out/widen_e2e.txt; : > $out
run() { label=$1; shift
  env "$@" python bench.py --steps 100 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$label e2e %.3e (%.3f ms)'%(d['e2e']['value'],d['e2e']['ms_per_step']))" >> $out
}
run "cached t16" FRC_WIDEN_NT=0
run "nt     t16" FRC_WIDEN_NT=1
run "cached t8 " FRC_WIDEN_NT=0 FRC_WIDEN_THREADS=8
run "nt     t8 " FRC_WIDEN_NT=1 FRC_WIDEN_THREADS=8
run "cached t12" FRC_WIDEN_NT=0 FRC_WIDEN_THREADS=12
run "nt     t12" FRC_WIDEN_NT=1 FRC_WIDEN_THREADS=12
run "cached t4 " FRC_WIDEN_NT=0 FRC_WIDEN_THREADS=4
run "cached t16" FRC_WIDEN_NT=0
run "nt     t16" FRC_WIDEN_NT=1
