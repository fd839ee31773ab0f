#!/bin/bash
# usage: tools/sass_stats.sh <binary> <function-substring>  -> loop structure + instruction mix of one kernel
BIN=$1; FN=$2
cuobjdump -sass $BIN | awk -v fn="$FN" '/Function :/{f=($0 ~ fn)} f' > /tmp/scratch/cur.sass
echo "total: $(grep -c -E '^\s+/\*[0-9a-f]{4,5}\*/' /tmp/scratch/cur.sass)"
grep -E "BRA|EXIT|CALL|RET" /tmp/scratch/cur.sass | sed 's/\/\* 0x[0-9a-f]* \*\///'
