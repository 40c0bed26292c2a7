#!/bin/bash
# Per-kernel resource usage of libsgs.so (registers, spills via STACK, shared memory).
cuobjdump --dump-resource-usage "$1" 2>/dev/null | grep -A1 "Function" | grep -v "^--" | paste - - | sed -e 's/ Function \(.*\):/\1/' | while read -r name rest; do
  echo "$(echo "$name" | c++filt | cut -c1-110) | $(echo "$rest" | grep -oE 'REG:[0-9]+|STACK:[0-9]+|SHARED:[0-9]+' | tr '\n' ' ')"
done
