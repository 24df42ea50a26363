#!/bin/sh
# Compiles the reference's delimiter FSM translation unit IN PLACE from /root/reference (never copied into
# this repo) together with oracle/ref_delim_driver.cpp into oracle/_ref/libref_delim.so.
# Only possible in the build container (the GPU box has no /root/reference; it uses the prebuilt .so).
set -e
cd "$(dirname "$0")"
REF=${LC_REFERENCE:-/root/reference}
[ -d "$REF/core/parser" ] || { echo "no reference checkout at $REF"; exit 0; }
mkdir -p _ref
g++ -O2 -std=c++17 -fPIC -shared -include cstring -I"$REF/core" -Ishim \
    "$REF/core/parser/DelimiterModeFsmParser.cpp" ref_delim_driver.cpp -o _ref/libref_delim.so
echo "built oracle/_ref/libref_delim.so"
