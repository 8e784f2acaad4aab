#!/bin/bash
# Build the library of another commit into fft-implementation-in-c_b200/lib/ab_<name>.so for same-box A/B timing:
#   tools/ab_build.sh <commit> <name>;  then on the GPU: FFTB200_LIB=fft-implementation-in-c_b200/lib/ab_<name>.so python tools/quick.py ...
set -e
commit=$1; name=$2
wt=/tmp/ab_wt_$name
rm -rf $wt; git worktree prune; git worktree add -f $wt $commit >/dev/null 2>&1
python $wt/fft-implementation-in-c_b200/build.py >/dev/null
cp $wt/fft-implementation-in-c_b200/lib/libfft_b200.so /root/repo/fft-implementation-in-c_b200/lib/ab_$name.so
git worktree remove --force $wt
echo built ab_$name.so from $commit
