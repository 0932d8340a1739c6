#!/bin/bash
# Same-box A/B of two builds of libditb200.so.  Box-to-box spread on this pool is +-3 % (the SM clock the power cap
# settles at differs per box), so numbers from two gpurun calls cannot be compared; this puts both variants into ONE.
#
#   tools/ab.sh build <name> [<git-ref>]     here (CPU): build the library of <git-ref> (default: the working tree)
#                                            into ab/<name>.so.  ab/ is git-ignored but travels with gpurun.
#                                            DITB200_NVCC_EXTRA=-DDITB200_PDL tools/ab.sh build pdl  builds a flag variant
#                                            (re-run `python -m fast_dit_b200.build` afterwards to restore the default library).
#   tools/ab.sh run <nameA> <nameB> -- <cmd> on the GPU box: run <cmd> with each library in turn, twice (A B A B),
#                                            e.g.  gpurun -- 'bash tools/ab.sh run old new -- python tools/tc_probe.py --no-cublas'
# `run` leaves <nameB> installed as fast_dit_b200/lib/libditb200.so on that (throw-away) box only.
set -e
cd "$(dirname "$0")/.."
LIB=fast_dit_b200/lib/libditb200.so
case "$1" in
  build)
    name=$2; ref=$3
    mkdir -p ab
    if [ -z "$ref" ]; then
      python -m fast_dit_b200.build > /dev/null && cp $LIB ab/$name.so
    else
      tmp=ab/_wt_$name; rm -rf "$tmp"; git worktree prune
      git worktree add --detach "$tmp" "$ref" > /dev/null
      (cd "$tmp" && python -m fast_dit_b200.build > /dev/null) && cp "$tmp/$LIB" ab/$name.so
      git worktree remove --force "$tmp"
    fi
    ls -la ab/$name.so ;;
  run)
    a=$2; b=$3; shift 4
    for v in $a $b $a $b; do
      cp ab/$v.so $LIB
      echo "== $v"
      timeout 600 "$@" 2>&1 | tail -4
    done ;;
  *) sed -n 2,12p "$0" ;;
esac
