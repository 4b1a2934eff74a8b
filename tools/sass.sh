#!/bin/bash
# usage: tools/sass.sh <object> <mangled-function-name>   -> opcode listing of one kernel
cuobjdump -sass -fun "$2" "$1" 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/\/\* 0x[0-9a-f]+ \*\///; s/^\s+\/\*[0-9a-f]+\*\/\s+//' 
