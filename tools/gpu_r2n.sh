#!/bin/bash
for st in 3 0 7; do
  echo "== STC_CONVH_STAGED=$st"
  STC_CONVH_STAGED=$st timeout 300 python tools/convh_prof.py 64 64 512 3 64 64 512 5 64 64 512 7 128 128 256 3 128 128 256 7 256 256 128 3 128 64 512 3 2>&1 | cut -c1-250
done
