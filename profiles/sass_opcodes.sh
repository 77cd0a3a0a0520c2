#!/bin/bash
# SASS evidence per object of libcasmvs.so (run after `make -C casmvsnet_pl_b200/csrc`):
# tcgen05 (UTCHMMA, LDTM/STTM, UTCBAR), TMA (UTMALDG, UBLKCP), mbarrier (SYNCS), packed fp32
# (FFMA2), 256-bit global accesses, local-memory traffic (spills).
cd "$(dirname "$0")/../casmvsnet_pl_b200/csrc/build" || exit 1
printf "%-22s %8s %6s %6s %7s %8s %7s %6s %6s %8s %8s %6s %6s\n" object UTCHMMA LDTM STTM UTCBAR UTMALDG UBLKCP SYNCS FFMA2 LDG.256 STG.256 LDL STL
for o in *.o; do
  s=$(cuobjdump -sass "$o")
  c() { echo "$s" | grep -c "$1"; }
  printf "%-22s %8d %6d %6d %7d %8d %7d %6d %6d %8d %8d %6d %6d\n" "$o" "$(c UTCHMMA)" "$(c 'LDTM')" "$(c 'STTM')" "$(c UTCBAR)" "$(c UTMALDG)" "$(c UBLKCP)" "$(c 'SYNCS')" "$(c FFMA2)" "$(c 'LDG.E.*256')" "$(c 'STG.E.*256')" "$(c ' LDL')" "$(c ' STL')"
done
