#!/bin/bash
# Where are the spills / barriers / reductions of a kernel?  tools/sass_phases.sh obj-or-so mangled-name-substring
f=$1; pat=$2
cuobjdump -sass "$f" | awk -v pat="$pat" '/Function :/ {on = index($0, pat) > 0} on' > /tmp/sass_one.txt
grep -c . /tmp/sass_one.txt
grep -n "LDL\|STL\|BAR.SYNC\|REDG\|RED\.\|BRA\b" /tmp/sass_one.txt | grep -v "BRA" | awk '{print $1, $2, $3, $4, $5, $6}' 
