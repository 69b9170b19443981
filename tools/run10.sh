#!/bin/bash
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x 2>&1 | tail -3
timeout 200 python tools/prof_conv.py l1_up l0_down l0_pair l3_3x3 | cut -c1-110
