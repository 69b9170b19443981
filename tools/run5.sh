#!/bin/bash
CASES="l0_pair l0_pair_res l0_pair_cat l0_pair_skip l0_pair_plain l0_3x3 l1_3x3 l1_pair l3_3x3 l3_geglu l0_kv l0_toout l0_1x1"
echo "== issuer = highest warp id"; timeout 300 python tools/prof_conv.py $CASES 2>&1 | cut -c1-45
echo "== issuer = warp 1 (round 1 layout)"; DAC_LIB=da-clip_b200/libdac_b200_dbg2.so timeout 300 python tools/prof_conv.py $CASES 2>&1 | cut -c1-45
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x 2>&1 | tail -2
