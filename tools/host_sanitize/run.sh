#!/bin/bash
# ThreadSanitizer and Address/UB-Sanitizer over the host sequence planner (paced, predicted walks). CPU only.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"; ROOT="$(dirname "$(dirname "$HERE")")"; OUT=${1:-/tmp/poppy_host_sanitize}
mkdir -p "$OUT"
SRC="$HERE/drive.cpp $ROOT/poppy_b200/csrc/host/host_abi.cpp $ROOT/poppy_b200/csrc/host/delaunay.cpp $HERE/stubs.cpp"
INC="-I $ROOT/include -I $ROOT/poppy_b200/csrc/host"
g++ -O1 -g -std=c++17 -fsanitize=thread -ffp-contract=off $INC $SRC -o "$OUT/drive_tsan" -pthread
g++ -O1 -g -std=c++17 -fsanitize=address,undefined -fno-sanitize-recover=undefined -ffp-contract=off $INC $SRC -o "$OUT/drive_asan" -pthread
"$OUT/drive_tsan" && "$OUT/drive_asan"
