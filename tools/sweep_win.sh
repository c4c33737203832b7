#!/bin/bash
# runs tools/quick_win.py on every build/variants/libspn_b200_<tag>.so given (GPU box)
nx=${NX:-64}
for tag in "$@"; do
  echo "=== variant $tag"
  SPN_B200_LIB=$PWD/build/variants/libspn_b200_$tag.so QUICK_FAST=1 timeout 300 python tools/quick_win.py $nx 2>&1 | grep -v "^mesh"
done
