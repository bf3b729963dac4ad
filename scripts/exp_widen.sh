#!/bin/bash
# e2e against how the host widens the fp32 bands (chunk size of the dynamic hand-out; 110000 = one static share per thread).
out=gpurun_out/widen_e2e.txt; : > $out
run() { label=$1; shift
  env "$@" python bench.py --steps 100 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$label e2e %.3e (%.3f ms)'%(d['e2e']['value'],d['e2e']['ms_per_step']))" >> $out
}
run "chunk 16384 " FRC_WIDEN_CHUNK=16384
run "chunk 110000" FRC_WIDEN_CHUNK=110000
run "chunk 4096  " FRC_WIDEN_CHUNK=4096
run "chunk 32768 " FRC_WIDEN_CHUNK=32768
run "chunk 16384 " FRC_WIDEN_CHUNK=16384
run "chunk 110000" FRC_WIDEN_CHUNK=110000
run "chunk 8192  " FRC_WIDEN_CHUNK=8192
