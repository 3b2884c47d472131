#!/bin/bash
# ring-depth sensitivity of the tensor-core convolution
run() { echo "== $*"; env "$@" NO_CUDNN=1 MVPNET_B200_DEBUG=1 timeout 120 python tools/conv_bench.py 2>&1 | grep -v "decoder1\|decoder2\|layer2\|decoder3" | sed 's/.*\(asets=[0-9]* stages=[0-9]* tps=[0-9]*\).*/   \1/' | cut -c1-78 | uniq; }
run MVPNET_B200_CONV_TPS=9 MVPNET_B200_CONV_ASETS=2 MVPNET_B200_CONV_STAGES=3
run MVPNET_B200_CONV_TPS=3 MVPNET_B200_CONV_ASETS=3
run MVPNET_B200_CONV_TPS=3 MVPNET_B200_CONV_ASETS=2
run MVPNET_B200_CONV_TPS=1 MVPNET_B200_CONV_ASETS=3
run MVPNET_B200_CONV_TM=2 MVPNET_B200_CONV_TPS=9
