#!/bin/bash
# Build libswgl_b200.so of another revision into tools/ab/<name>.so (development: A/B on one GPU box).
#   tools/ab_build.sh <git-rev> <name>
set -e
rev=$1; name=$2
root=$(cd "$(dirname "$0")/.." && pwd)
wt=/tmp/swgl_ab_$name
rm -rf "$wt"; git -C "$root" worktree prune; git -C "$root" worktree add -f --detach "$wt" "$rev" >/dev/null
(cd "$wt" && python -m swgl_b200.build >/dev/null)
mkdir -p "$root/tools/ab"; cp "$wt/swgl_b200/libswgl_b200.so" "$root/tools/ab/$name.so"
git -C "$root" worktree remove --force "$wt"
echo "$root/tools/ab/$name.so"
