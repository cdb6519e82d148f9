#!/usr/bin/env bash
# compute-sanitizer passes over a small selection of the GPU parity tests (memcheck + racecheck).
mkdir -p gpurun_out
SEL="test_image_parity or test_gradients_match_oracle_autograd or test_backward_generations_agree or test_sort_degenerate or test_rope_kernel_matches or test_rope_qk or test_fused_image_losses or test_colors_precomp or test_graph_replay_with_a_denser_scene or test_training_overflow or test_gradient_switches or test_fused_head or test_odd_sizes or test_raw_head or test_cuda_decoder_matches or test_pose_align"
for tool in memcheck racecheck; do
  timeout 2400 compute-sanitizer --tool $tool --error-exitcode 86 --log-file gpurun_out/sanitizer_$tool.log \
      python -m pytest tests -m gpu -q -k "$SEL" -p no:cacheprovider > gpurun_out/sanitizer_${tool}_pytest.log 2>&1
  echo "$tool exit=$?"
  tail -2 gpurun_out/sanitizer_${tool}_pytest.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/sanitizer_$tool.log | tail -2
  grep -c "=========     at " gpurun_out/sanitizer_$tool.log
done
