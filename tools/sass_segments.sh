#!/bin/bash
# static SASS instructions of the default step kernel between barriers: tools/sass_segments.sh lib.so
cuobjdump -sass $1 | awk '/Function : _Z15pve_step_kernelILi128ELi128ELi96ELb0/{f=1;next} /Function :/{f=0} f&&/^ +\/\*[0-9a-f]+\*\/ /{n++; seg++; if ($0 ~ /BAR\.|EXIT/) {printf "%d ", seg; seg=0}} END{print "| total", n}'
