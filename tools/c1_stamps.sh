#!/bin/bash
# Timeline of the fused conv_block1 kernel: stamped build -> tools/c1_stamps.py -> normal build again.
# Run from the repo root on a GPU box: bash tools/c1_stamps.sh > gpurun_out/c1_stamps.txt
F='-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr'
touch sound-event-detection_b200/csrc/sed_conv.cu
make -C sound-event-detection_b200/csrc NVFLAGS="$F -DSED_C1_STAMPS" > /dev/null 2>&1
timeout 120 python tools/c1_stamps.py
touch sound-event-detection_b200/csrc/sed_conv.cu
make -C sound-event-detection_b200/csrc > /dev/null 2>&1
