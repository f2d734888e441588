#!/bin/bash
for so in build_variants/*.so; do echo "== $so"; S2ST_B200_LIB=$PWD/$so timeout 200 python tools/sustained_check.py 2>&1 | sed -n 2p; done
timeout 200 python tools/sustained_check.py 2>&1 | sed -n 2p
S2ST_GL_PERSISTENT=0 timeout 200 python tools/sustained_check.py 2>&1 | sed -n 2p
