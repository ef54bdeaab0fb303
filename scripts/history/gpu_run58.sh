#!/bin/bash
cd /root/repo
for m in 32 16 8; do
  echo "micro_images=$m $(NQA_MICRO_IMAGES=$m timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | grep -o "\"value\": [0-9.]*\|sm_mhz\": [0-9.]*" | tr '\n' ' ')"
done
