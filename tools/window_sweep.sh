#!/bin/bash
# sweep of the window size of the large-m DMMA kernel (CA_FORCE_WINDOW_MODES) at one basis
for w in "$@"; do
  CA_FORCE_WINDOW_MODES=$w python tools/config5_multi_gpu.py 5,5,16 9472 2>/dev/null | tail -1 > gpurun_out/win_$w.json
  python -c "import json; d=json.load(open('gpurun_out/win_$w.json')); print('W', $w, 'rhs_ms', round(d['rhs_ms'],2), 'jvp_ms', round(d['jvp_ms'],2), 'fma/state', d['executed_fma_per_state'])"
done
