#!/bin/bash
# ThreadSanitizer + AddressSanitizer/UBSan builds of the native runtime pieces that own threads (CPU only; the CUDA kernels have
# their own compute-sanitizer script, bench/sanitize.sh).  Usage: bash bench/sanitize_native.sh [outdir]
set -u
root=$(cd "$(dirname "$0")/.." && pwd)
out=${1:-$root/profiles/sanitizer_native}
mkdir -p "$out"
tmp=$(mktemp -d)
src="$root/tests/native/data_loader_stress.cc $root/tepdist_b200/csrc/runtime/data_loader.cc"
rc=0
for san in thread "address,undefined"; do
  tag=${san%%,*}
  g++ -std=c++17 -O1 -g -fsanitize=$san -fno-omit-frame-pointer -I "$root/tepdist_b200/csrc" $src -o "$tmp/stress_$tag" -pthread || { rc=1; continue; }
  mkdir -p "$tmp/$tag"
  "$tmp/stress_$tag" "$tmp/$tag" > "$out/data_loader_$tag.log" 2>&1
  code=$?
  echo "[$tag] exit $code: $(tail -n 1 "$out/data_loader_$tag.log")"
  [ $code -ne 0 ] && rc=1
  grep -c "WARNING: ThreadSanitizer\|ERROR: AddressSanitizer\|runtime error:" "$out/data_loader_$tag.log" | sed "s/^/[$tag] sanitizer reports: /"
done
rm -rf "$tmp"
exit $rc
