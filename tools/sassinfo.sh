#!/bin/bash
# usage: tools/sassinfo.sh <mangled-name-substring>  -> /tmp/s2.txt + phase boundaries + opcode histogram
cuobjdump -sass cccl_b200/libb200rs.so 2>/dev/null | awk '/Function : /{f=$3} {print f"\t"$0}' | grep "$1" | awk -F'\t' '{print $2}' | grep -E '^\s+/\*[0-9a-f]{4}\*/' > /tmp/s2.txt
wc -l /tmp/s2.txt
awk '{n++; if ($0 ~ /BAR\.SYNC|EXIT/) print n": "$0}' /tmp/s2.txt | cut -c1-80
