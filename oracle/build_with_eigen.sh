#!/bin/sh
# Builds oracle/_ref/libeigen_sites.so: the reference's Eigen call sites on the hot path (plane fit QR, the two 23x23
# inverses of the gain, the 6x6 EigenSolver) behind a C ABI — see oracle/ref_eigen/eigen_sites.cpp.
# Needs Eigen3 (>= 3.3.4, include/IKFoM/README.md:16).  This image has none (no network either), so here the script only
# reports that; on a machine with Eigen:   sh oracle/build_with_eigen.sh [/path/to/eigen3]
# and then   python -m pytest tests/test_eigen_sites.py   pins the oracle's restatement against Eigen itself.
# The reference is built -O3 for baseline x86-64 (CMakeLists.txt:8,16): no -march=native / FMA here either.
set -e
cd "$(dirname "$0")"
INC="$1"
if [ -z "$INC" ]; then
  for d in /usr/include/eigen3 /usr/local/include/eigen3 /opt/homebrew/include/eigen3; do
    [ -f "$d/Eigen/Dense" ] && INC="$d" && break
  done
fi
if [ -z "$INC" ] || [ ! -f "$INC/Eigen/Dense" ]; then
  echo "build_with_eigen.sh: Eigen3 not found (pass its include directory); libeigen_sites.so not built" >&2
  exit 3
fi
mkdir -p _ref
g++ -std=c++14 -O3 -fPIC -shared -I"$INC" -o _ref/libeigen_sites.so ref_eigen/eigen_sites.cpp
echo "built oracle/_ref/libeigen_sites.so against $INC"
