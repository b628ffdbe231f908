#!/bin/bash
# Runs the GPU parity suite in isolated processes (a trapped kernel poisons its CUDA context, not the others).
# Usage on the GPU box:  bash scripts/gpu_check.sh   -> logs under gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() {  # name, timeout, pytest args...
  local name=$1; shift; local to=$1; shift
  timeout $to python -m pytest "$@" -q --timeout=300 -p no:cacheprovider > gpurun_out/$name.log 2>&1
  echo "== $name exit $? : $(tail -1 gpurun_out/$name.log)"
}
run e2i 600 tests/test_event2img_gpu.py -m gpu
run gemm 300 tests/test_encoder_gpu.py -m gpu -k "gemm_plain"
run gemm_epi 300 tests/test_encoder_gpu.py -m gpu -k "gemm_epilogues"
run ln 300 tests/test_encoder_gpu.py -m gpu -k "layernorm"
run attn 300 tests/test_encoder_gpu.py -m gpu -k "test_attention"
run enc 600 tests/test_encoder_gpu.py -m gpu -k "encoder"
run cls 600 tests/test_classifiers_gpu.py -m gpu
run train_kernels 600 tests/test_train_kernels_gpu.py -m gpu
run train 900 tests/test_train_gpu.py -m gpu
run tta 300 tests/test_tta.py -m gpu
run formats 300 tests/test_formats.py -m gpu
