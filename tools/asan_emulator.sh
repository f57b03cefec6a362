#!/bin/bash
# compute-sanitizer is closed on the GPU pool; the per-thread device functions are the same code the CPU emulator runs, so an
# AddressSanitizer + UBSan build of the emulator under its own test-suite is the bounds check that can be had:
#   tools/asan_emulator.sh [pytest -k expression]      (the full suite takes hours under ASan; the MSM + bucket-sort part ~7 min)
set -e
cd "$(dirname "$0")/.."
PKG=zksnap-circuits-halo2_b200
g++ -O1 -g -fsanitize=address,undefined -fno-sanitize-recover=undefined -std=c++17 -fPIC -shared -x c++ -I$PKG/csrc -I/usr/local/cuda/include -include cuda_runtime.h -o /tmp/libzkb200_hostemu_asan.so $PKG/csrc/hostemu.cu
ZKB200_EMU_LIB=/tmp/libzkb200_hostemu_asan.so LD_PRELOAD="$(g++ -print-file-name=libasan.so) $(g++ -print-file-name=libubsan.so)" ASAN_OPTIONS=detect_leaks=0 \
  python -m pytest tests/test_emulator.py -x -q -p no:cacheprovider ${1:+-k "$1"}
