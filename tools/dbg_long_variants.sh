# step time (panel 0 end / steps) and total forward time of the long-pair kernel variants
for v in "RSD_LONG_C=2 RSD_LONG_K=1" "RSD_LONG_C=2 RSD_LONG_K=2" "RSD_LONG_C=2 RSD_LONG_K=4" "RSD_LONG_C=4 RSD_LONG_K=1" "RSD_LONG_C=4 RSD_LONG_K=2" "RSD_LONG_C=4 RSD_LONG_K=4"; do
  echo "== $v"; env $v timeout 120 python tools/dbg_long.py 2>&1 | grep -E "forward ms|panel 0:|rror" | tail -3
done
