#!/bin/bash
# round-1 experiment: static vs dynamic schedule, persistent vs many-wave grids (64 GiB region)
S=""
for cps in 4 8 16 32 64 128; do S="$S,st256:$cps:256:4:1:0:1"; done          # static, more waves
for cps in 2 4; do for chunk in 65536 131072 262144 524288 1048576; do S="$S,st256:$cps:256:4:1:$chunk:2"; done; done
for cps in 1 2; do for chunk in 131072 262144 1048576; do S="$S,st256:$cps:512:4:1:$chunk:2,st256:$cps:1024:2:1:$chunk:2"; done; done
for cps in 2 4; do for chunk in 262144; do S="$S,st256:$cps:256:8:1:$chunk:2,st256:$cps:256:2:1:$chunk:2,st128:$cps:256:4:1:$chunk:2,st128:$cps:256:8:1:$chunk:2"; done; done
for cps in 1 2 4; do for grp in 1 2 4 8; do S="$S,tma:$cps:128:$grp:1:32768:2,tma:$cps:32:$grp:1:65536:2"; done; done
for cps in 2 4 8; do for chunk in 131072 262144 1048576; do S="$S,ld256:$cps:256:4:3:$chunk:2,ld256:$cps:256:2:3:$chunk:2"; done; done
S="memset,st256:8:256:4${S},ld256"
python benchmarks/profile_target.py --gib 64 --seq "$S" --reps 2 --warm 1
