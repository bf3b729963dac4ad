"""frackyfrac_b200 — a B200-native (sm_100a) UniFrac engine behind fluhus/frackyfrac's `frcfrc`.

  build      compiles the native pieces in-tree (python -m frackyfrac_b200.build)
  engine     ctypes binding of the C ABI (include/frcfrc_cuda.h): Context, Job, unifrac(), plan_bands()
  hostlib    ctypes binding of the C++ host side: Newick / table readers, species resolution, Go %v writer,
             compressed file IO, the sprspr converter
  dist       one process per GPU: communicator set-up, band ownership, merging the ranks' bands
  synth      the synthetic trees and tables of BASELINE.json's configurations

There is no CPU fallback: without libfrcfrc_cuda and an sm_100 device every compute call raises.
"""
