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
{
dump c2_wg_cube_f32_16x16x16       build/wg_cube.o _ZN4pfft14wg_cube_kernelIfLi16ELi1ELb0ELb1EEEvNS_8CubeArgsE
dump m512_wg_cube_f32_8x8x8        build/wg_cube.o _ZN4pfft14wg_cube_kernelIfLi8ELi4ELb0ELb1EEEvNS_8CubeArgsE
dump m2048_wg_rows3_f32_16x16x8    build/wg_cube.o _ZN4pfft15wg_rows3_kernelIfLi16ELi16ELi8ELi2ELb0EEEvNS_8CubeArgsE
dump d4096_wg_cube_f64_16x16x16    build/wg_cube.o _ZN4pfft14wg_cube_kernelIdLi16ELi1ELb0ELb1EEEvNS_8CubeArgsE
dump d2048_wg_rows3_f64_16x16x8    build/wg_cube.o _ZN4pfft15wg_rows3_kernelIdLi16ELi16ELi8ELi1ELb0EEEvNS_8CubeArgsE
dump l1d_wg_col_f32_256_inplace    build/wg_col.o _ZN4pfft13wg_col_kernelIfLi16ELi16ELi1ELi0ELb0ELb1EEEvNS_10PassParamsE14CUtensorMap_stb
dump l1d_wg_col_f32_256_rows       build/wg_col.o _ZN4pfft13wg_col_kernelIfLi16ELi16ELi1ELi2ELb0ELb0EEEvNS_10PassParamsE14CUtensorMap_stb
dump c4_wg_col_f64_256_inplace     build/wg_col.o _ZN4pfft13wg_col_kernelIdLi16ELi16ELi1ELi0ELb0ELb1EEEvNS_10PassParamsE14CUtensorMap_stb
dump c5_wg_col512_f32              build/wg_col.o _ZN4pfft16wg_col512_kernelENS_10PassParamsE14CUtensorMap_stb
dump s16_wi_tma_f32_16             build/wi_tma.o "$(cuobjdump -sass build/wi_tma.o | grep -oE '_ZN4pfft13wi_tma_kernelIfLi16E[A-Za-z0-9_]*' | head -1)"
dump fused_wg_fused2_f32           build/wg_fused.o _ZN4pfft16wg_fused2_kernelIfLi0EEEvNS_10PassParamsES1_14CUtensorMap_stNS_9FusedArgsEbb
} | tee $O/INDEX.txt
