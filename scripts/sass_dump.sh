#!/bin/bash
# Regenerates the SASS listings under profiles/ from the SHIPPED library (run on the CPU container after every kernel
# change: cuobjdump needs no GPU), with a one-line census of the instructions the design claims (TMA bulk copies and
# mbarrier operations in the hot kernel, packed fp32, 128-bit gathers).   usage: scripts/sass_dump.sh [tag]
cd "$(dirname "$0")/.."
tag=${1:-r02}
lib=monodepth2_b200/lib/libmd2loss.so
dump() {   # dump <mangled-substring-regex> <out>
  local sym
  sym=$(cuobjdump -elf "$lib" 2>/dev/null | grep -o "_ZN3md2[A-Za-z0-9_]*" | sort -u | grep -E "$1" | head -1)
  [ -z "$sym" ] && { echo "no symbol matches $1"; return 1; }
  cuobjdump -sass -fun "$sym" "$lib" 2>/dev/null | grep -v "^$" > "$2"
  echo "$2: $(c++filt "$sym" | cut -c1-110)"
  for op in UBLKCP UTMALDG SYNCS LDG.E.128 LDG.E.NA.128 LDGSTS FFMA2 FADD2 FMUL2 LDS.128 STS.128 BAR.SYNC LDL STL MUFU; do
    printf "    %-14s %s\n" "$op" "$(grep -c "[[:space:]]$op" "$2")"
  done
}
# the headline instantiation: role kernel, 2 sources, per-pixel minimum, automask, gradients, SSIM, packed A and B
dump "md2_march_rolesINS_6RoleOfINS_3CfgILi2ELb0ELb1ELb1ELb0EEEEELb1" profiles/${tag}_march_roles_2src_automask_grad.sass | tee profiles/${tag}_sass_census.txt
dump "md2_march_rolesINS_6RoleOfINS_3CfgILi3ELb0ELb1ELb1ELb0EEEEELb0" profiles/${tag}_march_roles_3src_automask_grad.sass | tee -a profiles/${tag}_sass_census.txt
dump "md2_identity_tmaILi2ELb0" profiles/${tag}_identity_tma_2src.sass | tee -a profiles/${tag}_sass_census.txt
python scripts/ptxas_summary.py monodepth2_b200/lib/libmd2loss.md2_kernels.cu.ptxas.log "" > profiles/${tag}_ptxas_registers.txt
echo "kernels with spills: $(grep -vc ' 0/   0 spill' profiles/${tag}_ptxas_registers.txt)" | tee -a profiles/${tag}_sass_census.txt
