#!/bin/bash
# usage: sass_hist.sh <binary-or-so> [function-substring]   - opcode histogram per function
cuobjdump -sass "$1" | awk -v pat="$2" '
/Function :/ {fn=$3}
/^ +\/\*[0-9a-f]+\*\/ / { if (pat=="" || index(fn,pat)) { op=$2; sub(/;$/,"",op); n=split(op,a,"."); base=a[1]; if (base=="IMAD" && n>1 && (a[2]=="WIDE"||a[2]=="MOV"||a[2]=="SHL"||a[2]=="IADD"||a[2]=="HI")) base=base"."a[2]; c[fn" "base]++; t[fn]++ } }
END { for (k in c) print c[k], k; for (f in t) print t[f], f, "TOTAL" }' | sort -k2,2 -k1,1nr
