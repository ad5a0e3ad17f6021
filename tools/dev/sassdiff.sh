#!/bin/bash
# usage: sassdiff.sh <git-rev> <file.cu>  — compile the file at <rev> and in the working tree without line info, compare SASS
set -e
REV=$1; F=$2; B=$(basename $F .cu)
rm -rf /tmp/th/sd && mkdir -p /tmp/th/sd/old/pcl_b200/csrc /tmp/th/sd/old/include /tmp/th/sd/new
git -C /root/repo archive $REV pcl_b200/csrc include | tar -x -C /tmp/th/sd/old
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -Xcompiler -fPIC"
(cd /tmp/th/sd/old && /usr/local/cuda/bin/nvcc $FLAGS -c pcl_b200/csrc/$B.cu -o /tmp/th/sd/old.o 2>/dev/null)
(/usr/local/cuda/bin/nvcc $FLAGS -c pcl_b200/csrc/$B.cu -o /tmp/th/sd/new.o 2>/dev/null)
/usr/local/cuda/bin/cuobjdump -sass /tmp/th/sd/old.o > /tmp/th/sd/old_full.sass
/usr/local/cuda/bin/cuobjdump -sass /tmp/th/sd/new.o > /tmp/th/sd/new_full.sass
python /root/repo/tools/dev/sasscmp.py /tmp/th/sd/old_full.sass /tmp/th/sd/new_full.sass | tail -6
