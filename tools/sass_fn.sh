#!/bin/bash
# SASS of one kernel of the product library, encodings stripped:  tools/sass_fn.sh <mangled-name-substring> [lib.so]
lib=${2:-verticut_b200/lib/libverticut_gpu.so}
cuobjdump -sass "$lib" | awk -v pat="$1" '/Function :/ {on = index($0, pat) > 0} on' | grep -v '^\s*/\* 0x' | sed 's/\/\* 0x[0-9a-f]* \*\///; s/ *$//'
