#!/bin/bash
# round-1 experiment: verify kernel shapes with the dynamic schedule (64 GiB region)
S="ld256"
for cps in 1 2 4 8; do for th in 128 256 512 1024; do
  if [ $((cps*th)) -gt 2048 ] || [ $((cps*th)) -lt 512 ]; then continue; fi
  for un in 2 4 8; do
    if [ $un -eq 8 ] && [ $((cps*th)) -gt 1024 ]; then continue; fi
    for chunk in 65536 131072 262144; do for pol in 1 3; do
      S="$S,ld256:$cps:$th:$un:$pol:$chunk:2"
    done; done
  done
done; done
for cps in 2 4 8; do for chunk in 131072 262144; do S="$S,ld128:$cps:256:4:3:$chunk:2,ld128:$cps:256:8:3:$chunk:2,ld128:$cps:512:4:3:$chunk:2"; done; done
python benchmarks/profile_target.py --gib 64 --seq "$S" --reps 2 --warm 1
