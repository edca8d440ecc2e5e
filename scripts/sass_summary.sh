#!/bin/bash
# SASS opcode histogram + resource usage of the hot kernels in the in-tree library (no GPU needed):
#   bash scripts/sass_summary.sh > profiles/sass_summary_rNN.txt
LIB=openfermion_b200/libofb200.so
echo "# resource usage (cuobjdump -res-usage), $(git rev-parse --short HEAD 2>/dev/null)"
cuobjdump -res-usage $LIB 2>/dev/null | grep -A1 -E "k_apply_tile|k_apply_multi|k_csc_tile|k_lanczos_step|k_sector_apply_fast" | grep -E "Function|REG" | sed 's/^ *//' | paste - - | sed -E 's/Function ([^:]*):.*REG:([0-9]+).*STACK:([0-9]+).*SHARED:([0-9]+).*/\1 REG=\2 STACK=\3 SHARED=\4/' | c++filt | sort
for pat in "k_apply_tileILb1ELi0ELb0ELb1E" "k_apply_tileILb1ELi0ELb0ELb0E" "k_apply_tileILb1ELi2ELb0ELb1E"; do
    echo
    echo "# opcode histogram of $(echo _ZN3ofb12${pat} | c++filt 2>/dev/null || echo $pat) (cuobjdump -sass)"
    cuobjdump -sass -fun "$(cuobjdump -elf $LIB 2>/dev/null | grep -o "_ZN3ofb12${pat}[A-Za-z0-9_]*" | head -1)" $LIB 2>/dev/null \
        | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9T]+\s+)?([A-Z0-9_]+).*/\2/' | sort | uniq -c | sort -rn | head -24
done
for pat in "k_csc_tileILb1ELb0EiE" "k_csc_tileILb0ELb0EiE"; do
    echo
    echo "# opcode histogram of $pat (cuobjdump -sass)"
    cuobjdump -sass -fun "$(cuobjdump -elf $LIB 2>/dev/null | grep -o "_ZN3ofb10${pat}[A-Za-z0-9_]*" | head -1)" $LIB 2>/dev/null \
        | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9T]+\s+)?([A-Z0-9_]+).*/\2/' | sort | uniq -c | sort -rn | head -16
done
