#!/bin/bash
# e2e (host buffers through the C ABI) against the number of output bands; FRC_WIRE=f64 = doubles on PCIe.
out=gpurun_out/bands.txt; : > $out
run() { # label, env...
  label=$1; shift
  env "$@" python bench.py --steps 100 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$label e2e %.3e (%.3f ms) value %.3e'%(d['e2e']['value'],d['e2e']['ms_per_step'],d['value']))" >> $out
}
run "bands 8 " FRC_BANDS=8
run "bands 8 " FRC_BANDS=8
run "bands 12" FRC_BANDS=12
run "bands 16" FRC_BANDS=16
run "bands 24" FRC_BANDS=24
run "bands 8 f64" FRC_BANDS=8 FRC_WIRE=f64
run "bands 8 " FRC_BANDS=8
