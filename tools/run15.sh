#!/bin/bash
# d = 32 attention: MUFU / polynomial split of the exponentials
for p in 0 1 2 3 4; do echo "== poly $p"; DAC_ATTN_POLY=$p python tools/prof_attn.py; done
