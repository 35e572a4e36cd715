mkdir -p gpurun_out
for tool in racecheck; do
  for cfg in "1024 4 2" "512 3 3" "2048 1 2" "512 3 6 overlapped" "256 2 3" "64 2 3"; do
    echo "== compute-sanitizer --tool $tool python scripts/san_target.py $cfg"
    timeout 600 compute-sanitizer --tool $tool --print-limit 6 python scripts/san_target.py $cfg 2>&1 | grep -E "SUMMARY|Error:|access at|checksums|hazards" | cut -c1-330 | head -14
  done
done > gpurun_out/r2x_racecheck.log 2>&1
echo done
