#!/bin/bash
timeout 200 python scratch/insitu_c2.py 2>&1 | grep -v "^$" | head -40
