#!/bin/bash
# timing experiments on the tensor-core convolution (MVPNET_B200_CONV_DBG bits: 1 one product of three, 2 epilogue
# without global memory, 4 weight ring without traffic, 16 no MMAs, 32 no patch traffic)
[ -n "$SKIP_TESTS" ] || timeout 300 python -m pytest tests/test_gpu_conv.py -x -q 2>&1 | tail -4
for dbg in ${DBGS:-0 1 16}; do
  echo "== DBG=$dbg"
  MVPNET_B200_CONV_DBG=$dbg NO_CUDNN=1 timeout 120 python tools/conv_bench.py 2>&1 | grep -v "^\[tc_conv" | cut -c1-80
done
for tm in ${TMS:-}; do
  echo "== TM=$tm"
  MVPNET_B200_CONV_TM=$tm NO_CUDNN=1 timeout 120 python tools/conv_bench.py 2>&1 | grep -v "^\[tc_conv" | cut -c1-80
done
