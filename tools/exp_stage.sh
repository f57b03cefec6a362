#!/bin/bash
python -m pytest tests/test_multi_device.py -m gpu -x -q 2>&1 | tail -3
for t in default 2 4 6 8 12 16; do if [ "$t" = default ]; then python tools/pageable_ab.py; else ZKB_STAGE_THREADS=$t python tools/pageable_ab.py; fi; done 2>&1 | grep stage_threads
tools/run_bench_n.sh 2 r2f_n2 2>&1 | grep "^rc=\|^witness_like\|^parity\|^value\|^single_process" | cut -c1-700
