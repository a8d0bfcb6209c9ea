#!/bin/bash
# opcode histogram of one kernel: sass_hist.sh file.o 'substring of mangled name' [top-n]
cuobjdump -sass "$1" | awk -v pat="$2" '/Function :/ {on = index($0, pat) > 0} on && /^ +\/\*[0-9a-f]+\*\/ / {op=$2; if (op ~ /^@/) op=$3; sub(/\..*/, "", op); gsub(/;/, "", op); h[op]++; n++} END {for (k in h) printf "%6d %s\n", h[k], k; printf "%6d TOTAL\n", n}' | sort -rn | head -${3:-30}
