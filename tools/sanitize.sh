#!/bin/bash
# compute-sanitizer over a reduced pass of every kernel family (run under gpurun; logs -> gpurun_out/sanitize_*.log, the
# summaries are committed under profiles/).  racecheck: shared-memory hazards (the staging rings, the column buffers);
# synccheck: named barriers / mbarriers / __syncthreads under divergence; memcheck: out-of-bounds and misaligned accesses.
O=gpurun_out
mkdir -p $O
python tools/sanitize_subset.py > $O/sanitize_plain.log 2>&1 || { echo "subset fails without the sanitizer"; tail -5 $O/sanitize_plain.log; exit 1; }
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_subset.py > $O/sanitize_$tool.log 2>&1
  echo "$tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize subset ok' $O/sanitize_$tool.log | tr '\n' ' ')"
done
