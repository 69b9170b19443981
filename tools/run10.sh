#!/bin/bash
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "stem" 2>&1 | tail -3
timeout 100 python tools/prof_conv.py stem stem_pair | cut -c1-100
timeout 600 python -m pytest tests/test_unet_gpu.py -q -x 2>&1 | tail -3
