#!/bin/bash
for cfg in "4 3" "8 3" "8 2" "12 2" "6 4"; do
  set -- $cfg
  echo "== BS=$1 EXTRA=$2"
  STC_CONVH_BS=$1 STC_CONVH_EXTRA=$2 timeout 300 python tools/convh_prof.py 64 64 512 3 64 64 512 7 128 128 256 3 128 128 256 7 256 256 128 3 2>&1 | cut -c1-120
done
