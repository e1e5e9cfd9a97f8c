#!/bin/bash
# usage: scripts/sass_fn.sh <object> <mangled-name substring>  -> compact SASS listing (address + instruction) of the first match
cuobjdump -sass "$1" 2>/dev/null | awk -v pat="$2" '/Function :/{f=(index($0,pat)>0 && !done); if(f) done=1} f' | awk '/Function :/{n++} n<2' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*([0-9a-f]{4})\*\/\s+/\1 /; s/\s*\/\*.*$//'
