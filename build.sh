#!/bin/bash
# Builds the product library (sm_100a only) and the CPU oracle. Used by __graft_entry__.build().
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=dusk-blindbidproof_b200/libbbp_b200.so
SRC=dusk-blindbidproof_b200/csrc
if [ ! -f $OUT ] || [ -n "$(find $SRC include -newer $OUT -type f | head -1)" ]; then
  $NVCC -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xptxas -v -Xcompiler -fPIC -shared \
    -o $OUT $SRC/bbp_capi.cu 2> build_ptxas.log || { tail -50 build_ptxas.log; exit 1; }
fi
# the Unix-socket server shell (plain C++ over the C ABI)
SRV=dusk-blindbidproof_b200/bbp-blindbid-server
if [ ! -f $SRV ] || [ $SRC/server/bbp_server.cpp -nt $SRV ] || [ $OUT -nt $SRV ]; then
  g++ -O2 -std=c++17 -pthread -o $SRV $SRC/server/bbp_server.cpp -Ldusk-blindbidproof_b200 -lbbp_b200 -Wl,-rpath,'$ORIGIN'
fi
LG=dusk-blindbidproof_b200/bbp-loadgen   # replays request frames against the server, one connection per request (tools/server_bench.py)
if [ ! -f $LG ] || [ $SRC/server/bbp_loadgen.cpp -nt $LG ]; then
  g++ -O2 -std=c++17 -pthread -o $LG $SRC/server/bbp_loadgen.cpp
fi
make -s -C oracle liboracle.so
