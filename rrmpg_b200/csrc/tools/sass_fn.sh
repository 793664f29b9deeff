#!/bin/bash
# usage: sass_fn.sh <object> <mangled-name-substring> -> instructions of that function, one per line (address opcode operands)
cuobjdump -sass "$1" | awk -v pat="$2" '/Function :/{on=index($0,pat)>0} on' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*([0-9a-f]{4})\*\/\s+/\1 /; s/\s*\/\*.*$//'
