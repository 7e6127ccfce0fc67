#!/usr/bin/env bash
run() { python tools/kernel_times.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], 'me', d['kernels_avg_ms'].get('me_sad_kernel'))"; }
echo "fused templated: $(run)"
echo "unfused templated: $(ICSP_ME_FUSED=0 run)"
echo "unfused no graphs: $(ICSP_ME_FUSED=0 ICSP_GRAPHS=0 run)"
echo "fused no graphs: $(ICSP_GRAPHS=0 run)"
mkdir -p /dev/shm/t && cd /dev/shm/t && python - <<'PY'
import sys; sys.path.insert(0, '/root/repo')
from icspcodec_b200 import synth
for i in range(16): synth.make_clip("highmotion", 300, 1000 + i).tofile(f"s{i:02d}_cif.yuv")
open("list.txt", "w").write("".join(f"s{i:02d}_cif.yuv\n" for i in range(16)))
PY
ICSPENC_TIMING=1 /root/repo/icspcodec_b200/host/icspenc -i s00_cif.yuv -n 300 -q 8 --intraPeriod 10 --quiet
ICSPENC_TIMING=1 /root/repo/icspcodec_b200/host/icspenc --batch list.txt -n 300 -q 8 --intraPeriod 10 --wave 8
ICSPENC_TIMING=1 /root/repo/icspcodec_b200/host/icspenc --batch list.txt -n 300 -q 8 --intraPeriod 10 --wave 8 --no-recon
