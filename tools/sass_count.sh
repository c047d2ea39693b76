#!/bin/bash
# usage: tools/sass_count.sh <object> <mangled-name-substring>: instruction mix of the matching kernels
obj=$1; pat=$2
cuobjdump -sass "$obj" | awk -v pat="$pat" '
/Function :/ { name=$3; on = (index(name, pat) > 0); if (on) { print "== " name; } next }
on && /^ +\/\*[0-9a-f]+\*\// { op=$2; sub(/\..*/, "", op); if (op ~ /^@/) { op=$3; sub(/\..*/, "", op);} cnt[name" "op]++; tot[name]++ }
END { for (k in cnt) print cnt[k], k; for (n in tot) print tot[n], n, "TOTAL" }' | sort -k2,2 -k1,1nr | awk '{print}' 
