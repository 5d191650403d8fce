#!/bin/bash
# sub-band geometry sweep of srcnn_infer_rows_host on C3 (each configuration in its own process:
# the switches are read once)
cd "$(dirname "$0")/.."
for sb in 6 8 10 12 16 20 24; do
  for cap in 2 4 8; do
    echo -n "subbands $sb rampcap $cap: "
    SRCNN_E2E_SUBBANDS=$sb SRCNN_E2E_RAMPCAP=$cap python tools/e2e_infer.py | tail -1
  done
done
echo -n "graph off, 16/4: "; SRCNN_E2E_GRAPH=0 python tools/e2e_infer.py | tail -1
tools/probe/mn16_probe | tail -4
