#!/bin/bash
# Build a variant of libdmayolo.so that differs in ONE source file's -D flags (for same-box A/B through DMAY_SO):
#   tools/build_variant.sh nms "-DDMAY_FILTER_FLAT_PASS2=1" dma_yolo_b200/libdmayolo_v1.so
set -e
SRC=$1; DEFS=$2; OUT=$3
cd "$(dirname "$0")/.."
python -c "import sys; sys.path.insert(0, '.'); from dma_yolo_b200 import build; build.build()"
NV=/usr/local/cuda/bin/nvcc
$NV -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --expt-relaxed-constexpr -Xcompiler -fPIC $DEFS \
    -c dma_yolo_b200/csrc/$SRC.cu -o /tmp/variant_$SRC.o
OBJS=$(ls dma_yolo_b200/build/*.o | grep -v "/$SRC.o")
$NV -shared -o $OUT $OBJS /tmp/variant_$SRC.o -gencode arch=compute_100a,code=sm_100a
echo built $OUT
