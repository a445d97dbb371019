for l in 4 5 6; do for G in 2 8; do
  r=$(AMGB200_STREAM_G=$G timeout 100 python tools/prof_level.py p3d 128 $l 2>&1 | grep "ms per" | tail -1)
  echo "L$l G=$G $r"
done; done
