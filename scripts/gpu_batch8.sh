#!/usr/bin/env bash
cd scripts/probes
./tma_probe2
BOXW=32 BOXH=16 ./tma_probe2
BOXW=16 BOXH=16 ./tma_probe2
BOXW=64 BOXH=8 IMGW=128 IMGH=128 X=0 Y=0 ./tma_probe2
BOXW=32 BOXH=32 IMGW=1920 IMGH=1080 X=32 Y=32 ./tma_probe2
F32=1 ./tma_probe2
DLSYM=1 ./tma_probe2
CTAGROUP=1 ./tma_probe2
L2P=2 BOXW=32 BOXH=16 ./tma_probe2
X=0 Y=0 ./tma_probe2
X=32 Y=16 BOXW=32 BOXH=16 IMGW=128 IMGH=64 ./tma_probe2
