#!/bin/bash
SEGS_DECODE_VARIANT=2 timeout 200 python tools/timeline_decode.py 2>&1 | grep -E "span|decode_"
