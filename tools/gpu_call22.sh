#!/usr/bin/env bash
run() { python tools/kernel_times.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], 'me', d['kernels_avg_ms'].get('me_sad_kernel'), 'fdct', d['kernels_avg_ms'].get('fdct_quant_kernel'), 'idct', d['kernels_avg_ms'].get('idct_recon_kernel<enc>'), 'chain', d['kernels_avg_ms'].get('dc_chain_kernel'))"; }
echo "current (raw chunk, fused): $(run)"
echo "current, unfused: $(ICSP_ME_FUSED=0 run)"
cp icspcodec_b200/libicspcuda.so /tmp/keep.so; cp tmp_variants/me_rawchunk0.so icspcodec_b200/libicspcuda.so
echo "rawchunk0, fused: $(run)"
echo "rawchunk0, unfused: $(ICSP_ME_FUSED=0 run)"
echo "rawchunk0, unfused, no graphs: $(ICSP_ME_FUSED=0 ICSP_GRAPHS=0 run)"
cp /tmp/keep.so icspcodec_b200/libicspcuda.so
