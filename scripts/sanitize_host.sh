#!/bin/bash
# CPU-only memory-safety pass: builds the C++ host library and the C oracle with ASan + UBSan into their _build
# directories, runs the host / fuzz / oracle tests against them, then restores the regular builds.
# (compute-sanitizer is closed on the GPU pool, so this covers the host half; the CUDA half rests on the parity suites.)
set -e
cd "$(dirname "$0")/.."
H=frackyfrac_b200/_build/libfrcfrc_host.so; O=oracle/_build/liboracle.so
cp $H /tmp/host.bak; cp $O /tmp/oracle.bak
trap 'cp /tmp/host.bak $H; cp /tmp/oracle.bak $O' EXIT
SAN="-O1 -g -fPIC -pthread -fsanitize=address,undefined -fno-omit-frame-pointer"
g++ $SAN -std=c++17 -I include -shared -o $H frackyfrac_b200/host/hostlib.cpp frackyfrac_b200/host/fileio.cpp -lz -ldl
gcc $SAN -std=gnu11 -ffp-contract=off -shared -o $O oracle/unifrac_oracle.c -lm
ASAN_OPTIONS=detect_leaks=0 LD_PRELOAD=$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so) \
  python -m pytest tests/test_host.py tests/test_host_fuzz.py tests/test_oracle_kat.py -q -p no:cacheprovider 2>&1 | tee /tmp/sanitize.log | tail -3
if grep -q "runtime error\|AddressSanitizer" /tmp/sanitize.log; then echo "SANITIZER FINDINGS in /tmp/sanitize.log"; exit 1; fi
echo "clean"
