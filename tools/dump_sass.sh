#!/bin/bash
# SASS listings of the hot kernels (cuobjdump -sass of the in-tree libibt.so) -> profiles/sass/
set -e
cd "$(dirname "$0")/.."
SO=iceberg_tracking_code_b200/libibt.so
dump() { cuobjdump -sass "$SO" | awk -v pat="$1" '/Function : /{f = ($0 ~ pat)} f' | grep -vE "^\s+/\* 0x[0-9a-f]+ \*/\s*$" | sed -E "s@\s*/\* 0x[0-9a-f]+ \*/\s*\$@@" > "profiles/sass/$2.sass"; wc -l "profiles/sass/$2.sass"; }
dump "lk_kernelILi31ELi31" lk_kernel_31x31
dump "pyr_level_kernelILb1ELb1ELi4" pyr_level_kernel_deriv_down_4warps
dump "gray_c3_vec_kernelILi15" gray_c3_vec_kernel
dump "pyr_level_tma_kernelILb1ELb1ELi4" pyr_level_tma_kernel_deriv_down_4warps
dump "eig_nms_kernelILi10ELb0" eig_nms_kernel_bs10
dump "gftt_select_kernel" gftt_select_kernel
dump "eig_kernel" eig_kernel
dump "jpg_sync_round" jpg_sync_round
dump "jpg_idct" jpg_idct
dump "jpg_colorILi15ELi3ELi2ELi2" jpg_color
cuobjdump -sass "$SO" | grep -oE "^\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9T]+\s+)?[A-Z0-9_.]+" | awk '{print $NF}' | grep -E "^(IDP|LDGSTS|REDUX|UTMA|UBLKCP|HMMA|UTC|SYNCS|MATCH)" | sort | uniq -c | sort -rn > profiles/sass/opcode_evidence.txt
cat profiles/sass/opcode_evidence.txt
