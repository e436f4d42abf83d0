#!/bin/sh
# Builds the DEBUG-ONLY emulation harness (see emul_engine.cpp).  Not part of the product build.
set -e
cd "$(dirname "$0")/../.."
mkdir -p tests/emul/_build
g++ -std=c++17 -O1 -g -ffp-contract=off -Wall -Wno-unused-function -o tests/emul/_build/b200samtools_emul \
    samtools_b200/csrc/host/cli.cpp samtools_b200/csrc/host/hts_io.cpp tests/emul/emul_engine.cpp -lz
# the htslib-compatible iterator tier (plp_compat.cpp) + its test client, on the emulation harness
g++ -std=c++17 -O1 -g -ffp-contract=off -Wall -Wno-unused-function -Iinclude -o tests/emul/_build/plp_dump_emul \
    tests/compat/plp_dump.cpp samtools_b200/csrc/host/plp_compat.cpp samtools_b200/csrc/host/hts_io.cpp tests/emul/emul_engine.cpp -lz
