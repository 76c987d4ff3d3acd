#!/bin/bash
# usage: tools/sass_fn.sh <substring of mangled kernel name> [lib]  -> SASS of that kernel
LIB=${2:-ntjoin_b200/libmxe.so}
cuobjdump -sass "$LIB" | awk -v pat="$1" '/Function :/ {on = index($0, pat) > 0} on {print}'
