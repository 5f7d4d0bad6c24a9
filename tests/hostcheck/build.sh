#!/bin/sh
# TEST INFRASTRUCTURE: builds tests/hostcheck/libhostcheck.so (g++ only, links the host layer for its readers)
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
CXX=/usr/bin/g++; [ -x $CXX ] || CXX=g++
$CXX -std=c++17 -O2 -fPIC -shared -ffp-contract=fast -o "$HERE/libhostcheck.so" "$HERE/hostcheck.cpp" \
  -L"$HERE/../../is3d_b200" -lis3d_host -lis3d_b200 -Wl,-rpath,"$HERE/../../is3d_b200"
