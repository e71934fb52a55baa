# A/B of kuafu_b200/lib against kuafu_b200/lib_head on the other recipes
for sc in spheres active cornell articulated; do
  for lib in lib_head lib; do
    echo "== $sc $lib"; KFRT_LIB_DIR=kuafu_b200/$lib python tools/counters.py $sc 0 0 ${SPP:-16} 2>&1 | tail -2
  done
done
