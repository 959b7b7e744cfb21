#!/bin/bash
# compute-sanitizer over the cluster / DSMEM kernels (frame feedback, LSTM sampler + teacher, Adafactor big2d), the
# tcgen05 tap GEMMs incl. the fused CE epilogue and the ones-MMA bias gradient, and the latent block -- at the small
# shapes of the unit tests.  ONE tool per gpurun call (B200_PROFILING.md):  bash tools/sanitize.sh memcheck|racecheck
tool=${1:-memcheck}
mkdir -p gpurun_out
sel='tests/test_ops.py::test_frame_feedback_shift_median_noise_and_scheduled_sampling
tests/test_ops.py::test_teacher_forced_lstm_cluster_kernels_match_lstmcell
tests/test_ops.py::test_gru_blend
tests/test_ops.py::test_standardize_vs_oracle
tests/test_adafactor.py::test_fused_adafactor_matches_oracle
tests/test_model_parity.py::test_latent_sampling_bit_exact_stage_isolated
tests/test_tap_gemm.py::test_conv3x3
tests/test_tap_gemm.py::test_wgrad
tests/test_tap_gemm.py::test_wgrad_two_m_tiles_per_cta_matches_one
tests/test_latent.py::test_fused_latent_block_forward_and_backward_match_autograd[4-3]
tests/test_logits_ce.py::test_fused_logits_ce_matches_gemm_then_pixel_ce[2-False]'
timeout 1500 compute-sanitizer --tool $tool --error-exitcode 99 --print-limit 20 \
    python -m pytest $sel -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/sanitizer_$tool.log 2>&1
echo "compute-sanitizer $tool rc=$?" | tee -a gpurun_out/sanitizer_$tool.log
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error:|hazard" gpurun_out/sanitizer_$tool.log | tail -15
