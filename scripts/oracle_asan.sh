#!/bin/bash
# C oracle under AddressSanitizer + UBSan (CPU only): scripts/oracle_asan.sh [cases] [seed]
set -e
cd "$(dirname "$0")/.."
T=$(mktemp -d)
gcc -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -std=gnu99 -Iinclude -Ioracle -Ivideo-encoder_b200/host -o $T/oracle_asan \
    scripts/oracle_asan.c oracle/b2o_*.c video-encoder_b200/host/b2h_cavlc.c video-encoder_b200/host/b2h_cabac.c -lm -lpthread
ASAN_OPTIONS=detect_leaks=1 $T/oracle_asan "${1:-60}" "${2:-1}"
rm -rf $T
