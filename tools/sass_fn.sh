#!/bin/bash
# usage: tools/sass_fn.sh <regex on the (mangled) function name> [so]   -- prints the SASS of the first matching kernel
SO=${2:-rio_rs_b200/librio_cuda.so}
cuobjdump -sass "$SO" | awk -v pat="$1" '/Function :/ {on = ($0 ~ pat)} on {print}'
