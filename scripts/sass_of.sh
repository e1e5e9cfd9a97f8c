#!/bin/bash
# usage: scripts/sass_of.sh <object|so> <mangled-name-substring>  -> SASS of the first matching function
cuobjdump -sass "$1" | awk -v pat="$2" '/Function :/ {on = index($0, pat) > 0 && !done; if (on) done=1} on {print}'
