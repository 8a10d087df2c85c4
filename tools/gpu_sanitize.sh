# GPU-box job (gpurun): compute-sanitizer (memcheck / racecheck / synccheck / initcheck) over tools/sanitize_small.py
mkdir -p gpurun_out
: > gpurun_out/sanitizer.txt
for tool in memcheck racecheck synccheck initcheck; do
  echo "== $tool" >> gpurun_out/sanitizer.txt
  timeout 1200 compute-sanitizer --tool $tool python tools/sanitize_small.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_small ok|Error|error|hazard" | head -12 >> gpurun_out/sanitizer.txt
done
cat gpurun_out/sanitizer.txt
