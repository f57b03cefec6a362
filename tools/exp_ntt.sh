#!/bin/bash
echo "--- default (4 CTAs/SM)"; python tools/ntt_time.py
for v in 5 6; do echo "--- min CTAs $v"; ZKB200_LIB=$PWD/zksnap-circuits-halo2_b200/libzkb200_ntt$v.so python tools/ntt_time.py --cases 22x16 24x4; done
