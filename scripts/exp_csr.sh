#!/bin/bash
# frc_create's table staging: chunks per worker share (1 = static shares) against e2e and the share times.
out=gpurun_out/csr_e2e.txt; : > $out
run() { label=$1; shift
  env "$@" python bench.py --steps 100 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$label e2e %.3e (%.3f ms)'%(d['e2e']['value'],d['e2e']['ms_per_step']))" >> $out
  env "$@" FRC_TRACE=1 python bench.py --steps 2 --warmup 3 --no-cpu 2>&1 >/dev/null | grep "CSR share times\|joined" | tail -2 >> $out
}
run "chunks/share 8" FRC_CSR_CHUNKS=8
run "chunks/share 1" FRC_CSR_CHUNKS=1
run "chunks/share 2" FRC_CSR_CHUNKS=2
run "chunks/share 4" FRC_CSR_CHUNKS=4
run "chunks/share 8" FRC_CSR_CHUNKS=8
run "chunks/share 1" FRC_CSR_CHUNKS=1
