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
# The experimental CTA-pair variant (tcgen05 cta_group::2, off by default) runs separately: racecheck reports "potential
# RAW (CUDA barrier operation)" at shared offsets 0x58-0x5f of both CTAs for the `tcgen05.alloc.cta_group::2` instruction
# itself — inside the 1 KB system-reserved region below the user window's 0x400 base, i.e. the allocator's own handshake.
for tool in $TOOLS; do
  extra=""
  [ "$tool" = "racecheck" ] && extra="--racecheck-report all"
  timeout 1500 compute-sanitizer --tool "$tool" $extra --error-exitcode 86 --launch-timeout 0 \
      python -m pytest $SEL -k "not pair2" -q -x -p no:cacheprovider > "gpurun_out/sanitizer_${tool}.log" 2>&1
  echo "$tool exit code $?" >> "gpurun_out/sanitizer_${tool}.log"
  tail -n 12 "gpurun_out/sanitizer_${tool}.log"
  timeout 900 compute-sanitizer --tool "$tool" $extra --error-exitcode 86 --launch-timeout 0 \
      python -m pytest tests/test_gpu_kernels.py -k "pair2" -q -x -p no:cacheprovider > "gpurun_out/sanitizer_${tool}_pair2.log" 2>&1
  echo "$tool (pair2) exit code $?" >> "gpurun_out/sanitizer_${tool}_pair2.log"
  grep "SUMMARY\|exit code" "gpurun_out/sanitizer_${tool}_pair2.log"
done
