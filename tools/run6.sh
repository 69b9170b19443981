#!/bin/bash
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "kv_in_kernel or kv_tensor_core" 2>&1 | tail -8
timeout 300 python tools/prof_conv.py l0_kvtc l0_kvtc_pn l1_kvtc l1_kvtc_pn l3_geglu 2>&1 | cut -c1-60
