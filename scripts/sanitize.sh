#!/bin/bash
# compute-sanitizer passes over small GPU tests that exercise every hand-written mbarrier / TMEM / cluster protocol
# (tcgen05 GEMMs, strip-sweep convs, wgrad, training step, scan preparation, post-processing).  usage: scripts/sanitize.sh OUTDIR
OUT=${1:-gpurun_out}
mkdir -p $OUT
SEL='tests/test_gpu_gemm.py::test_dense_layer_backends tests/test_gpu_forward.py::test_forward_edge_inputs tests/test_gpu_forward.py::test_dense_volume_vs_oracle tests/test_gpu_forward.py::test_dense_candidate_compaction tests/test_gpu_train.py::test_train_step_matches_autograd_oracle tests/test_gpu_gather.py tests/test_gpu_prep.py::test_candidate_mask_and_bbox tests/test_gpu_postproc.py::test_quirks_absent_class_and_no_overlap tests/test_gpu_prep.py::test_upload_volume_box_writes_exactly_the_box tests/test_gpu_forward.py::test_sparse_mask_sweep_item_skipping'
for tool in ${TOOLS:-memcheck racecheck synccheck}; do
  start=$(date +%s)
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python -m pytest $SEL -x -q -k "not 4097 and not 333" > $OUT/sanitizer_$tool.log 2>&1
  rc=$?
  echo "== $tool rc=$rc $(( $(date +%s) - start )) s" | tee -a $OUT/sanitizer_summary.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" $OUT/sanitizer_$tool.log | tail -5 | tee -a $OUT/sanitizer_summary.txt
done
