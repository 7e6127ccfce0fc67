#!/usr/bin/env bash
# TEST INFRASTRUCTURE — builds the UNMODIFIED reference (JawThrow/ICSPCodec) from the sources where
# they lie under /root/reference into oracle/_ref/ (git-ignored, travels to the GPU box with gpurun).
# Nothing from the reference is copied into the repo.  Recipe follows SURVEY.md §8c:
#   canonical flags -O2, never -march=native / -ffast-math (FMA contraction changes the output).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${ICSP_REFERENCE_ROOT:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/source/encoder" ]; then
  echo "build_ref: $REF not present (GPU box?) - keeping prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
ENC="$REF/source/encoder"
DEC="$REF/source/decoder"
CXX="${CXX:-g++}"
# stock encoder CLI (-O2 canonical and -O3 "best CPU")
$CXX -O2 -w -pthread -o "$OUT/ICSPCodec_O2" "$ENC/encoder_main.cpp" "$ENC/ICSP_Codec_Encoder_source.cpp" "$ENC/ICSPCodec.cpp" "$ENC/ICSP_thread.cpp"
$CXX -O3 -w -pthread -o "$OUT/ICSPCodec_O3" "$ENC/encoder_main.cpp" "$ENC/ICSP_Codec_Encoder_source.cpp" "$ENC/ICSPCodec.cpp" "$ENC/ICSP_thread.cpp"
# stock decoder CLI (MSVC-era code: needs four forced includes)
$CXX -O2 -w -include cstring -include cmath -include cstdlib -include cstdio -o "$OUT/ICSPDecoder_O2" "$DEC/decode.cpp" "$DEC/ICSP_Codec_Decoder_source.cpp"
# function-level tap harness: our harness TU + the three non-main encoder TUs
$CXX -O2 -w -pthread -I"$ENC" -o "$OUT/ref_taps" "$HERE/ref_taps.cpp" "$ENC/ICSP_Codec_Encoder_source.cpp" "$ENC/ICSPCodec.cpp" "$ENC/ICSP_thread.cpp"
echo "build_ref: built $(ls "$OUT" | tr '\n' ' ')"
# reference-side binding (integration/icsp_ref_shim.cpp): the reference's own main / loader / makebitstream / dump on top of
# libicspcuda.  The reference's single_thread_encoding is renamed inside its own translation unit only (-D), every other
# reference source is compiled as it is; the shim provides the replacement definition.
ROOT="$(cd "$HERE/.." && pwd)"
if [ -f "$ROOT/icspcodec_b200/libicspcuda.so" ]; then
  $CXX -O2 -w -c -Dsingle_thread_encoding=ref_single_thread_encoding -o "$OUT/enc_source_renamed.o" "$ENC/ICSP_Codec_Encoder_source.cpp"
  $CXX -O2 -w -pthread -I"$ENC" -I"$ROOT/include" -o "$OUT/ICSPCodec_gpu" "$ENC/encoder_main.cpp" "$ENC/ICSPCodec.cpp" "$ENC/ICSP_thread.cpp" \
      "$OUT/enc_source_renamed.o" "$ROOT/integration/icsp_ref_shim.cpp" -L"$ROOT/icspcodec_b200" -licspcuda \
      -Wl,-rpath,'$ORIGIN/../../icspcodec_b200' -Wl,-rpath,/root/repo/icspcodec_b200
  rm -f "$OUT/enc_source_renamed.o"
  echo "build_ref: built ICSPCodec_gpu (reference front end + writer on libicspcuda)"
else
  echo "build_ref: libicspcuda.so not built yet - skipping ICSPCodec_gpu" >&2
fi
