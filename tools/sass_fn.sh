#!/bin/bash
# usage: tools/sass_fn.sh <mangled-substring> [so]  -> /tmp/fn.txt (one instruction per line)
SO=${2:-allwave_b200/liballwave_cuda.so}
cuobjdump -sass $SO | awk -v pat="$1" '/Function :/{on=index($0,pat)>0} on' | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" | awk '{ $NF=""; print }' | sed 's#/\* 0x[0-9a-f]* \*/##; s#/\*$##' > /tmp/fn.txt
wc -l /tmp/fn.txt
