#!/bin/bash
# Memory-safety check of the kernel SOURCES without a GPU (compute-sanitizer is closed on the B200 pool):
# builds the CPU emulation (tests/emu/) with AddressSanitizer and runs forward/backward/grouped cases through
# it.  Device buffers and shared memory are heap allocations there, so any out-of-bounds index aborts.
#   bash tools/asan_emu.sh
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=/tmp/admmtv_asan; mkdir -p $OUT; cd $OUT
F="-std=c++20 -O1 -g -fsanitize=address -fno-omit-frame-pointer -fPIC -pthread -DADMMTV_EMU -I$ROOT/tests/emu -I$ROOT/admm_deconv_b200/csrc -x c++"
g++ $F -c $ROOT/admm_deconv_b200/csrc/admmtv_api.cu -o api.o &
g++ $F -c $ROOT/admm_deconv_b200/csrc/loss_api.cu -o loss.o &
g++ $F -c $ROOT/admm_deconv_b200/csrc/inst_generic.cu -o generic.o &
g++ $F -c $ROOT/admm_deconv_b200/csrc/inst_small.cu -o small.o &
for l in 5 6 7 8 9 10 11 12 20 21 22 23 24 25 26 27 28 29 30 31; do
  S=""; case $l in 8|10|11|12|2[1-9]|3[01]) S="-DADMMTV_STUB";; esac
  g++ $F $S -DADMMTV_INST=$l -c $ROOT/admm_deconv_b200/csrc/inst_dim1.cu -o d1_$l.o &
  g++ $F $S -DADMMTV_INST=$l -c $ROOT/admm_deconv_b200/csrc/inst_dim2.cu -o d2_$l.o &
done
g++ -std=c++20 -O1 -g -fsanitize=address -fPIC -pthread -I$ROOT/tests/emu -c $ROOT/tests/emu/cuda_emu.cpp -o emu.o
wait
g++ -shared -fsanitize=address -pthread *.o -o libadmmtv_asan.so
ASAN_OPTIONS=detect_leaks=0 LD_PRELOAD=$(gcc -print-file-name=libasan.so) python $ROOT/tools/asan_cases.py $OUT/libadmmtv_asan.so
