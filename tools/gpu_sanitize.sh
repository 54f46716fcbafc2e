#!/bin/bash
# compute-sanitizer passes over the kernels that stage through shared memory (finalize family, fused word masks) and a
# few accumulate cases. Slow: keep the selections small.
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
echo "== memcheck: finalize family"; timeout 900 $CS --tool memcheck --error-exitcode 1 python -m pytest tests/test_finalize_gpu.py -q -x -k "golden or fast_and_generic or expand_words or short_prompts" 2>&1 | tail -6 | tee gpurun_out/sanitize_memcheck_finalize.txt
echo "== racecheck: finalize family"; timeout 900 $CS --tool racecheck --error-exitcode 1 python -m pytest tests/test_finalize_gpu.py -q -x -k "golden or fast_and_generic or expand_words_fused" 2>&1 | tail -6 | tee gpurun_out/sanitize_racecheck_finalize.txt
echo "== memcheck: accumulate (SIMT, probs, tcgen05 golden layers)"; timeout 900 $CS --tool memcheck --error-exitcode 1 python -m pytest tests/test_parity_elementwise_gpu.py tests/test_accumulate_gpu.py -q -x -k "golden_layers or materialised or weighted_partition or partial_and_tiny" 2>&1 | tail -6 | tee gpurun_out/sanitize_memcheck_accumulate.txt
