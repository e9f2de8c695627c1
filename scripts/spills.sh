#!/bin/bash
# usage: scripts/spills.sh <mangled kernel name>  -> where the LDL/STL of the run-ahead kernel sit
f=/tmp/ra_sass_$$.txt
cuobjdump -sass -fun "$1" v2e2v_b200/libv2e2v_b200.so 2>/dev/null | grep -v "^\s*/\* 0x" | sed 's/\/\* 0x[0-9a-f]* \*\///' | cut -c1-100 > $f
wc -l < $f
B=$(grep -n "BAR.SYNC" $f | head -1 | cut -d: -f1); E=$(grep -n "ATOMS.POPC.INC" $f | head -1 | cut -d: -f1)
LOOP=$(awk -v b=$B -v e=$E 'NR>=b && NR<=e' $f | grep -n "LDG.E.*CONSTANT\|LDG.E.128\|LDG.E.U8\|LDG.E.64" | head -3 | tr '\n' ' ')
echo "prologue-sync line $B, stream_cnt atomic line $E"
echo "LDL/STL lines: $(grep -n 'LDL\|STL' $f | awk -F: '{print $1}' | tr '\n' ' ')"
echo "CALL lines: $(grep -n 'CALL' $f | awk -F: '{print $1}' | tr '\n' ' ')"
grep -n "BAR.SYNC" $f | awk -F: '{print $1}' | tr '\n' ' '; echo
cp $f /tmp/ra_sass.txt
