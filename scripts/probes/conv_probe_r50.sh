#!/bin/bash
# where the time of the epilogue-bound 1x1 launches goes: ResNet50 ops 1 (downsample), 4 (layer1.0.conv3), 14 (layer2.0.conv3)
for m in 0 1 4; do
  echo "== CONV_PROBE=$m"
  for op in 1 4 14; do PDFUSION_B200_CONV_PROBE=$m timeout 120 python scripts/probes/stem_time.py $op 1 resnet50; done
done
