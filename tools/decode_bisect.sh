#!/bin/bash
# Bisect the decode program kernel's time with QEFT_DECODE_DEBUG (results are wrong when set): 1 = consumers skip the math,
# 2 = producer skips the scale / outlier cp.async, 4 = producer skips the bulk copies.  Prints one JSON line per setting.
for dbg in 0 1 2 3 4 5 6 7; do
  for slots in 4 2; do
    echo -n "{\"debug\": $dbg, \"slots\": $slots, \"result\": "
    QEFT_DECODE_DEBUG=$dbg QEFT_DECODE_SLOTS=$slots timeout 60 python tools/decode_program_time.py 7b 2>/dev/null | tr -d '\n'
    echo "}"
  done
done
