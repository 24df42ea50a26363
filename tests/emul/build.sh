#!/bin/sh
# builds the TEST-ONLY emulation library (see lc_emul.cpp header)
set -e
cd "$(dirname "$0")"
g++ -O2 -std=c++17 -fPIC -shared -Wall -o liblc_emul.so lc_emul.cpp ../../loongcollector_b200/csrc/regex_compiler.cpp
