#!/bin/bash
# On the GPU box: device time of the actor kernel with every .ab/lib*.so (or the names given).
NAMES=${*:-$(ls .ab/lib*.so | sed 's#.ab/lib##; s#\.so##')}
for n in $NAMES; do
  echo "== $n: $(PVE_MCC_LIBRARY=$PWD/.ab/lib$n.so timeout 300 python tools/actor_timing.py 2>&1 | grep 'actor kernel' | tail -1)"
done
