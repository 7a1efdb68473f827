#!/bin/bash
# round-1 experiment: per-WARP dynamic grabs (schedule 3) vs per-CTA (schedule 2), full arena
S="auto,vauto"
for shape in "1:512:4" "1:1024:4" "1:1024:2" "2:512:4" "1:256:4" "1:128:8"; do
  for chunk in 16384 32768 65536 131072; do
    S="$S,st256:$shape:1:$chunk:3"
  done
done
for shape in "1:1024:4" "1:1024:2" "2:512:4" "1:512:4" "1:1024:8"; do
  for chunk in 16384 32768 65536 131072; do
    S="$S,ld256:$shape:3:$chunk:3"
  done
done
python benchmarks/profile_target.py --gib 0 --seq "$S" --reps 3 --warm 1
