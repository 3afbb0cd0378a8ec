#!/bin/bash
python scripts/dbg_stamps.py 2>&1 | tail -20
python scripts/dbg_stamps.py SUPERKMER_TARGET=1000000000 2>&1 | tail -20
