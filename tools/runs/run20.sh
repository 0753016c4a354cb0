#!/bin/bash
for d in 0 1 2 4 6; do PASSL_B200_EPI_DEBUG=$d timeout 120 python tools/epi_probe.py 2>&1 | tail -1; done
