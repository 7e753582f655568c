#!/bin/bash
# builds a tuning variant of the CUDA library: tools/build_variant.sh <name> <extra nvcc flags...>  -> decipher_b200/csrc/build_<name>/libdecipher_b200.so
set -e
name=$1; shift
cd "$(dirname "$0")/../decipher_b200/csrc"
make BUILD=build_$name EXTRA="$*" build_$name/libdecipher_b200.so > /dev/null
grep -A2 "Compiling entry function '_Z21k_physics_resident_v2ILi4ELi[0-9]ELb0E" build_$name/dcp_resident_v2.ptxas.log | grep -v "^--" | tail -2
