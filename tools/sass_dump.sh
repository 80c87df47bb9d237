#!/bin/bash
# Full SASS listings of the kernels that carry the headline numbers (cuobjdump -sass on the objects `make` built for
# sm_100a) -> profiles/sass/<name>.sass, plus the TMA / mbarrier mnemonic counts per listing.
set -eu
O=profiles/sass; mkdir -p $O
dump() {  # name object mangled-function
  cuobjdump -sass -fun "$3" "$2" > $O/$1.sass
  printf "%-28s %6d instructions  UBLKCP=%d UTMALDG=%d UTMASTG=%d SYNCS=%d LDGSTS=%d BAR=%d LDL/STL=%d\n" "$1" \
    "$(grep -cE '^\s+/\*[0-9a-f]{4}\*/' $O/$1.sass)" "$(grep -c UBLKCP $O/$1.sass || true)" "$(grep -c UTMALDG $O/$1.sass || true)" \
    "$(grep -c UTMASTG $O/$1.sass || true)" "$(grep -c SYNCS $O/$1.sass || true)" "$(grep -c LDGSTS $O/$1.sass || true)" \
    "$(grep -cE '\bBAR\.' $O/$1.sass || true)" "$(grep -cE '\b(LDL|STL)\b' $O/$1.sass || true)"
}
fn() {  # object regex -> first mangled kernel name matching
  cuobjdump -sass "$1" | grep -oE "Function : [A-Za-z0-9_]+" | sed 's/Function : //' | grep -E "$2" | head -1
}
{
dump c2_wg_cube_f32_16x16x16       build/wg_cube.o "$(fn build/wg_cube.o 'wg_cube_kernelIfLi16ELi1ELb0ELb1ELi0E')"
dump m512_wg_cube_f32_8x8x8        build/wg_cube.o "$(fn build/wg_cube.o 'wg_cube_kernelIfLi8ELi4ELb0ELb1ELi0E')"
dump m2048_wg_rows3_f32_16x16x8    build/wg_cube.o "$(fn build/wg_cube.o 'wg_rows3_kernelIfLi16ELi16ELi8ELi2ELb0ELi1ELi0E')"
dump m8192_wg_rows3_f32_16x16x32   build/wg_cube.o "$(fn build/wg_cube.o 'wg_rows3_kernelIfLi16ELi16ELi32ELi1ELb0ELi2ELi0ELb1E')"
dump d4096_wg_cube_f64_16x16x16    build/wg_cube.o "$(fn build/wg_cube.o 'wg_cube_kernelIdLi16ELi1ELb0ELb1ELi0ELb1E')"
dump d2048_wg_rows3_f64_16x16x8    build/wg_cube.o "$(fn build/wg_cube.o 'wg_rows3_kernelIdLi16ELi16ELi8ELi1ELb0ELi1ELi0E')"
dump r2c8192_wg_cube_f32           build/wg_cube.o "$(fn build/wg_cube.o 'wg_cube_kernelIfLi16ELi1ELb0ELb1ELi1E')"
dump c2r8192_wg_cube_f32           build/wg_cube.o "$(fn build/wg_cube.o 'wg_cube_kernelIfLi16ELi1ELb0ELb1ELi2E')"
dump r2c512_wg_rows3_f32_16x4x4    build/wg_cube.o "$(fn build/wg_cube.o 'wg_rows3_kernelIfLi16ELi4ELi4ELi8ELb0ELi1ELi1E')"
dump l1d_wg_col_f32_256_inplace    build/wg_col.o "$(fn build/wg_col.o 'wg_col_kernelIfLi16ELi16ELi1ELi0ELb0ELb1E')"
dump l1d_wg_col_f32_256_rows       build/wg_col.o "$(fn build/wg_col.o 'wg_col_kernelIfLi16ELi16ELi1ELi2ELb0ELb0E')"
dump c4_wg_col_f64_256_inplace     build/wg_col.o "$(fn build/wg_col.o 'wg_col_kernelIdLi16ELi16ELi1ELi0ELb0ELb1E')"
dump c5_wg_col512_f32              build/wg_col.o "$(fn build/wg_col.o 'wg_col512_kernel')"
dump c3b_wg_colr3_tma_f32_split    build/wg_colr3.o "$(fn build/wg_colr3.o 'wg_colr3_tma_kernelIfLi10ELi10ELi10ELb0ELb0E')"
dump r2c32_wi_tma_real_f32         build/wi_tma.o "$(fn build/wi_tma.o 'wi_tma_real_kernelIfLi16ELi1E')"
dump s16_wi_tma_f32_16             build/wi_tma.o "$(fn build/wi_tma.o 'wi_tma_kernelIfLi16E')"
dump fused_wg_fused2_f32           build/wg_fused.o "$(fn build/wg_fused.o 'wg_fused2_kernelIfLi0E')"
} | tee $O/INDEX.txt
