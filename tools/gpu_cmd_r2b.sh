cd tools/micro
for a in "6016 128" "6016 512" "2048 128" "2048 512" "1024 128"; do
  echo "== upd_bench (engine shape) $a"; timeout 60 ./upd_bench $a 2>&1 | head -4
  echo "== upd3_bench $a"; timeout 60 ./upd3_bench $a 2>&1
done
