#!/bin/bash
# compute-sanitizer passes over the CUDA path: memcheck over the tiny end-to-end inference + training step
# (every kernel family incl. tcgen05 GEMMs, fused attention, decode / NMS), the GEMM format / tile sweep and the fused XLNet
# attention; racecheck (shared-memory hazards) over the end-to-end smoke.  Logs -> gpurun_out/r2_sanitizer_*.log
cd "$(dirname "$0")/.."
CS=${CS:-/usr/local/cuda/bin/compute-sanitizer}
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name: $*"; timeout 1500 "$@" > gpurun_out/r2_sanitizer_$name.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error:|hazard" gpurun_out/r2_sanitizer_$name.log | sort | uniq -c | head -8; }
run memcheck_smoke $CS --tool memcheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()"
run memcheck_train $CS --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_train.py -q -m gpu -k "test_model_gradients_vs_oracle_autograd or test_training_steps_flat" -x
run memcheck_gemm $CS --tool memcheck --error-exitcode 3 python tools/gemm2_probe.py check
XL_CHECK_ONLY=1 run memcheck_xlattn $CS --tool memcheck --error-exitcode 3 python tools/xl_probe.py mixed
run memcheck_round2_kernels $CS --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_model.py tests/test_gpu_zz_nlq.py -q -m gpu -x -k "window_attention or single_pass_self or fpn1d_vs_reference or nlq_detections"
run racecheck_round2_kernels $CS --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_model.py -q -m gpu -x -k "window_attention_kernel and 1024 or single_pass_self and 128 or fpn1d_vs_reference"
run racecheck_smoke $CS --tool racecheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()"
