#!/bin/bash
# compute-sanitizer passes over the kernels of the hot path (small shapes): memcheck, racecheck, synccheck, initcheck.
# Usage (GPU box): bash tools/sanitize.sh [outdir]   -> <outdir>/r2_sanitizer_<tool>.txt
out=${1:-gpurun_out}
sel='tc_scorer_layers or tc3_split_scorer or fused_features or features_multi or topk or kernel_merge or head_tensor_core or features_without_side or mask_early_out or boxes_to_mask or pack_poses or sharded_prefilter'
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 --log-file $out/r2_sanitizer_$tool.txt \
      python -m pytest tests -m gpu -q -x -k "$sel" > $out/r2_sanitizer_${tool}_pytest.txt 2>&1
  echo "== $tool rc=$? : $(tail -1 $out/r2_sanitizer_${tool}_pytest.txt)"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error:|hazard" $out/r2_sanitizer_$tool.txt | sort | uniq -c | head -8
done
