#!/bin/bash
# compute-sanitizer over the kernel unit tests and a tiny end-to-end synthesis (run on a GPU box through gpurun):
#   memcheck  — out-of-bounds / misaligned global, shared and TMEM-adjacent accesses
#   racecheck — shared-memory hazards between the producer / MMA / epilogue roles (the hand-rolled mbarrier pipelines)
# Logs land in gpurun_out/ and are summarised into profiles/ by hand.  Usage: tools/sanitize.sh [memcheck|racecheck|synccheck ...]
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TOOLS="${*:-memcheck racecheck}"
SEL='tests/test_gpu_kernels.py tests/test_gpu_parity.py::test_golden_fixture tests/test_gpu_parity.py::test_cfg1_short_utterance tests/test_gpu_bert.py::test_golden_fixture'
for tool in $TOOLS; do
  extra=""
  [ "$tool" = "racecheck" ] && extra="--racecheck-report all"
  timeout 1500 compute-sanitizer --tool "$tool" $extra --error-exitcode 86 --launch-timeout 0 \
      python -m pytest $SEL -q -x -p no:cacheprovider > "gpurun_out/sanitizer_${tool}.log" 2>&1
  echo "$tool exit code $?" >> "gpurun_out/sanitizer_${tool}.log"
  tail -n 12 "gpurun_out/sanitizer_${tool}.log"
done
